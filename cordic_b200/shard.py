"""Sharding of a sample stream over the GPUs of one box.

Samples are independent, so a stream of n samples is cut into `world` contiguous ranges, one per rank, with
every boundary a multiple of `align` samples (128 = one warp-block of the seeded kernel, and 512 bytes of
phase words, so every shard keeps the 16-byte alignment the vector kernels want).  There is no data-path
collective: a rank needs nothing from its peers.  For the NCO a rank does not even need an input slice: its
first phase follows from (phase0, step, start) in closed form.
"""


def shard_range(n, world, rank, align=128):
    """(start, count) of rank's contiguous share of n samples; the last rank takes the ragged tail."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = (n // world) // align * align
    if per == 0:                       # tiny stream: rank 0 takes everything
        return (0, n) if rank == 0 else (n, 0)
    start = rank * per
    count = per if rank < world - 1 else n - start
    return start, count


def nco_start_phase(phase0, step, start):
    """32-bit accumulator value at sample index `start`: phase0 + start*step (mod 2^32)."""
    return (phase0 + start * step) & 0xFFFFFFFF
