/* cordic_b200/vshim/numer_fft.c -- the two symbols bench/cpp/fft.h:46-47 declares, for the SFDR print of
 * cordic_tb.cpp:342-373 (printed, never asserted).  The reference links FFTW3 here (bench/cpp/fftw.c);
 * FFTW is not in this image, so this is a plain iterative radix-2 FFT in double precision. */
#include <assert.h>
#include <math.h>
#include <stdlib.h>

unsigned nextlg(unsigned long vl) {
	unsigned long r = 1;
	assert(vl > 0);
	while (r < vl) r <<= 1;
	return (unsigned)r;
}

static void bitrev_permute(double *d, unsigned long n) {
	unsigned long j = 0;
	for (unsigned long i = 0; i < n; i++) {
		if (i < j) {
			double a = d[2 * i], b = d[2 * i + 1];
			d[2 * i] = d[2 * j]; d[2 * i + 1] = d[2 * j + 1];
			d[2 * j] = a; d[2 * j + 1] = b;
		}
		unsigned long bit = n >> 1;
		for (; bit && (j & bit); bit >>= 1) j ^= bit;
		j |= bit;
	}
}

void numer_fft(double *data, unsigned nn, int isign) {
	const unsigned long n = nn;
	const double dir = (isign < 0) ? -1.0 : 1.0;
	double *tw = (double *)malloc(sizeof(double) * n);	/* n/2 complex twiddles */
	assert(tw);
	for (unsigned long k = 0; k < n / 2; k++) {
		double a = dir * 2.0 * M_PI * (double)k / (double)n;
		tw[2 * k] = cos(a); tw[2 * k + 1] = sin(a);
	}
	bitrev_permute(data, n);
	for (unsigned long span = 1; span < n; span <<= 1) {
		const unsigned long step = n / (span << 1);
		for (unsigned long base = 0; base < n; base += span << 1) {
			for (unsigned long k = 0; k < span; k++) {
				const double wr = tw[2 * k * step], wi = tw[2 * k * step + 1];
				double *p = data + 2 * (base + k), *q = data + 2 * (base + k + span);
				const double tr = q[0] * wr - q[1] * wi, ti = q[0] * wi + q[1] * wr;
				q[0] = p[0] - tr; q[1] = p[1] - ti;
				p[0] += tr; p[1] += ti;
			}
		}
	}
	free(tw);
}
