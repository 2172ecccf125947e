/* oracle/shim/numer_fft.c -- TEST INFRASTRUCTURE ONLY.
 * Stands in for bench/cpp/fftw.c (which needs libfftw3, absent from this image): the
 * two symbols bench/cpp/fft.h:46-47 declares.  Used only for the SFDR figure that
 * cordic_tb.cpp:342-373 prints and never asserts.  Iterative radix-2 DIT FFT with
 * per-stage twiddles from sin/cos (double), forward for isign<0 like FFTW_FORWARD.
 */
#include <assert.h>
#include <math.h>
#include <stdlib.h>

unsigned nextlg(unsigned long vl) {
	unsigned long r;
	assert(vl > 0);
	for (r = 1; r < vl; r <<= 1)
		;
	return (unsigned)r;
}

void numer_fft(double *data, unsigned nn, int isign) {
	unsigned long n = nn, i, j, k, m;
	/* bit reversal */
	for (i = 0, j = 0; i < n; i++) {
		if (i < j) {
			double tr = data[2 * i], ti = data[2 * i + 1];
			data[2 * i] = data[2 * j]; data[2 * i + 1] = data[2 * j + 1];
			data[2 * j] = tr; data[2 * j + 1] = ti;
		}
		m = n >> 1;
		while (m >= 1 && (j & m)) { j ^= m; m >>= 1; }
		j |= m;
	}
	double *wr = (double *)malloc(sizeof(double) * (n / 2 + 1));
	double *wi = (double *)malloc(sizeof(double) * (n / 2 + 1));
	assert(wr && wi);
	for (k = 0; k < n / 2; k++) {
		double a = (isign < 0 ? -2.0 : 2.0) * M_PI * (double)k / (double)n;
		wr[k] = cos(a); wi[k] = sin(a);
	}
	for (unsigned long len = 2; len <= n; len <<= 1) {
		unsigned long half = len >> 1, stride = n / len;
		for (i = 0; i < n; i += len) {
			for (k = 0; k < half; k++) {
				double cr = wr[k * stride], ci = wi[k * stride];
				double *a = &data[2 * (i + k)], *b = &data[2 * (i + k + half)];
				double tr = b[0] * cr - b[1] * ci, ti = b[0] * ci + b[1] * cr;
				b[0] = a[0] - tr; b[1] = a[1] - ti;
				a[0] += tr; a[1] += ti;
			}
		}
	}
	free(wr); free(wi);
}
