// oracle/shim/verilated_vcd_c.h -- TEST INFRASTRUCTURE ONLY.
// A small value-change-dump writer with the VerilatedVcdC surface the reference's
// TESTB uses (bench/cpp/testb.h:67-81,96-105): open / dump / flush / close.  Models
// register their signals through declare(); dump() does change detection over all of
// them and flush() pushes the buffer to the file, which is the per-tick cost profile
// of the real thing.  Set ZC_SHIM_NOTRACE=1 in the environment to make open() a no-op
// (the "VCD off" leg of the CPU baseline).
#ifndef ZC_SHIM_VERILATED_VCD_C_H
#define ZC_SHIM_VERILATED_VCD_C_H

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

class VerilatedVcdC {
	struct Sig { std::string name; int width; const uint32_t *ptr; uint32_t last; bool first; };
	std::vector<Sig> m_sigs;
	FILE *m_fp;
	bool m_header_done;
	static void code(unsigned k, char *out) {
		int n = 0;
		do { out[n++] = (char)(33 + (k % 94)); k /= 94; } while (k);
		out[n] = 0;
	}
public:
	VerilatedVcdC() : m_fp(NULL), m_header_done(false) {}
	~VerilatedVcdC() { close(); }
	void declare(const char *name, int width, const uint32_t *ptr) {
		Sig s; s.name = name; s.width = width; s.ptr = ptr; s.last = 0; s.first = true;
		m_sigs.push_back(s);
	}
	void open(const char *fname) {
		const char *off = getenv("ZC_SHIM_NOTRACE");
		if (off && off[0] == '1')
			return;
		m_fp = fopen(fname, "w");
	}
	void dump(uint64_t t) {
		if (!m_fp) return;
		char id[8];
		if (!m_header_done) {
			fprintf(m_fp, "$version zcordic oracle shim $end\n$timescale 1ns $end\n"
				"$scope module TOP $end\n");
			for (unsigned k = 0; k < m_sigs.size(); k++) {
				code(k, id);
				fprintf(m_fp, "$var wire %d %s %s $end\n", m_sigs[k].width, id,
					m_sigs[k].name.c_str());
			}
			fprintf(m_fp, "$upscope $end\n$enddefinitions $end\n");
			m_header_done = true;
		}
		fprintf(m_fp, "#%lu\n", (unsigned long)t);
		for (unsigned k = 0; k < m_sigs.size(); k++) {
			Sig &s = m_sigs[k];
			uint32_t v = *s.ptr;
			if (!s.first && v == s.last) continue;
			s.first = false; s.last = v;
			code(k, id);
			if (s.width == 1) {
				fprintf(m_fp, "%c%s\n", (v & 1) ? '1' : '0', id);
			} else {
				char bits[40]; int n = 0;
				bits[n++] = 'b';
				for (int b = s.width - 1; b >= 0; b--)
					bits[n++] = ((v >> b) & 1) ? '1' : '0';
				bits[n] = 0;
				fprintf(m_fp, "%s %s\n", bits, id);
			}
		}
	}
	void flush() { if (m_fp) fflush(m_fp); }
	void close() { if (m_fp) { fclose(m_fp); m_fp = NULL; } }
};

#endif
