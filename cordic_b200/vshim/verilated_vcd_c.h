// cordic_b200/vshim/verilated_vcd_c.h -- VerilatedVcdC surface used by TESTB (testb.h:67-81,96-105).
// The batched engine has no per-clock internal state to dump, so tracing is accepted and ignored.
#ifndef ZC_VSHIM_VERILATED_VCD_C_H
#define ZC_VSHIM_VERILATED_VCD_C_H
#include <cstdint>
class VerilatedVcdC {
public:
	void open(const char *) {}
	void dump(uint64_t) {}
	void flush() {}
	void close() {}
};
#endif
