// oracle/shim/verilated.h -- TEST INFRASTRUCTURE ONLY.
// Minimal stand-in for the Verilator runtime header so that the reference's own,
// unmodified bench/cpp/cordic_tb.cpp and bench/cpp/topolar_tb.cpp compile in an image
// that has no Verilator.  Only the surface those files use is provided
// (bench/cpp/testb.h:49-136, bench/cpp/cordic_tb.cpp:88, bench/cpp/topolar_tb.cpp:91).
#ifndef ZC_SHIM_VERILATED_H
#define ZC_SHIM_VERILATED_H

#include <cassert>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

typedef uint8_t  CData;
typedef uint16_t SData;
typedef uint32_t IData;
typedef uint64_t QData;

class Verilated {
public:
	static void commandArgs(int, char **) {}
	static void traceEverOn(bool) {}
};

#endif
