// cordic_b200/vshim/verilated.h -- the slice of the Verilator runtime API that the reference's test benches
// touch (bench/cpp/testb.h:49-136, cordic_tb.cpp:88, topolar_tb.cpp:91), so that those sources compile,
// unmodified, against the GPU-backed models in this directory.
#ifndef ZC_VSHIM_VERILATED_H
#define ZC_VSHIM_VERILATED_H
#include <cassert>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
class Verilated {
public:
	static void commandArgs(int, char **) {}
	static void traceEverOn(bool) {}
};
#endif
