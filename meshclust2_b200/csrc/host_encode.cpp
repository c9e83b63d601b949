// host_encode.cpp — host side of the input contract for callers that do not already hold a
// reference ChromosomeOneDigit: raw DNA text -> (one-digit codes, inclusive segment list, effective size).
//
// Behaviour follows Chromosome::help (src/nonltr/Chromosome.cpp:130-154): upper-case, split at N
// (removeAmbiguous :263-291, including its quirk that a run opening on the very last base is dropped), bridge gaps
// of < 10 and drop merged segments < 20 bp when the sequence is longer than 20 (mergeSegments :298-353), cut
// segments > 1 Mbp (makeSegmentList :355-385), then ChromosomeOneDigit::encode
// (src/nonltr/ChromosomeOneDigit.cpp:79-133) with the DNA code map (src/nonltr/ChromosomeOneDigitDna.cpp:48-68).
// It is plain host C++ (table driven, one classification pass) and stays on the host exactly as in the reference.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../../include/meshclust2_b200.h"

namespace mc2 {
void set_error(const std::string &msg);
}

namespace {

struct Lut {
	unsigned char up[256];   // upper-cased byte
	signed char code[256];   // one-digit code of the upper-cased byte, -1 if not a nucleotide letter
	Lut()
	{
		for (int c = 0; c < 256; c++) {
			up[c] = (unsigned char)((c >= 'a' && c <= 'z') ? c - 32 : c);
			code[c] = -1;
		}
		const char *letters = "ACGTRYMKSWHBVDNX";
		const signed char val[] = {0, 1, 2, 3, 2, 1, 0, 3, 2, 3, 1, 3, 0, 3, 1, 2};
		for (int i = 0; letters[i]; i++) {
			code[(unsigned char)letters[i]] = val[i];
		}
	}
};
const Lut g_lut;

struct Seg {
	int32_t s, e;
};

int encode_one(const char *text, uint64_t len, char *out, std::vector<Seg> &segs, uint64_t *eff)
{
	segs.clear();
	*eff = 0;
	if (len >= (1ULL << 31)) {
		return MC2_ERR_ARG;
	}
	const int64_t n = (int64_t)len;
	// runs of non-N
	std::vector<Seg> runs;
	int64_t i = 0;
	while (i < n) {
		while (i < n && g_lut.up[(unsigned char)text[i]] == 'N') {
			i++;
		}
		if (i >= n) {
			break;
		}
		int64_t s = i;
		while (i < n && g_lut.up[(unsigned char)text[i]] != 'N') {
			i++;
		}
		if (s == n - 1) {
			break; // a run opening on the last base is never closed (Chromosome.cpp:268-285)
		}
		runs.push_back(Seg{(int32_t)s, (int32_t)(i - 1)});
	}
	// bridge + drop short
	std::vector<Seg> merged;
	if (n > 20) {
		if (!runs.empty()) {
			Seg cur = runs[0];
			for (size_t j = 1; j < runs.size(); j++) {
				if (runs[j].s - cur.e < 10) {
					cur.e = runs[j].e;
				} else {
					if (cur.e - cur.s + 1 >= 20) {
						merged.push_back(cur);
					}
					cur = runs[j];
				}
			}
			if (cur.e - cur.s + 1 >= 20) {
				merged.push_back(cur);
			}
		}
	} else {
		merged = runs;
	}
	// 1 Mbp fragments; the last fragment takes the remainder
	const int32_t kFrag = 1000000;
	for (const Seg &m : merged) {
		int32_t L = m.e - m.s + 1;
		if (L > kFrag) {
			int32_t nf = L / kFrag;
			for (int32_t h = 0; h < nf; h++) {
				int32_t fs = m.s + h * kFrag;
				segs.push_back(Seg{fs, h == nf - 1 ? m.e : fs + kFrag - 1});
			}
		} else {
			segs.push_back(m);
		}
	}
	// encode: everything is upper-cased; letters become codes except N outside segments (kept as 'N');
	// a non-letter is an error inside a segment, and outside one only when at least one segment exists
	size_t sj = 0;
	const bool have = !segs.empty();
	for (i = 0; i < n; i++) {
		while (sj < segs.size() && segs[sj].e < i) {
			sj++;
		}
		const bool in = sj < segs.size() && segs[sj].s <= i;
		unsigned char u = g_lut.up[(unsigned char)text[i]];
		signed char c = g_lut.code[u];
		if (in) {
			if (c < 0) {
				return MC2_ERR_INPUT;
			}
			out[i] = (char)c;
		} else if (!have || u == 'N') {
			out[i] = (char)u;
		} else {
			if (c < 0) {
				return MC2_ERR_INPUT;
			}
			out[i] = (char)c;
		}
	}
	uint64_t e = 0;
	for (const Seg &s : segs) {
		e += (uint64_t)(s.e - s.s + 1);
	}
	*eff = e;
	return MC2_OK;
}

} // namespace

extern "C" {

int mc2_encode_dna(const char *text, uint64_t len, char *codes_out, int32_t *segs_out, uint64_t max_segs, uint64_t *n_segs,
		   uint64_t *effective_size)
{
	if ((!text && len) || (!codes_out && len) || !n_segs || !effective_size) {
		mc2::set_error("mc2_encode_dna: NULL argument");
		return MC2_ERR_ARG;
	}
	std::vector<Seg> segs;
	int rc = encode_one(text, len, codes_out, segs, effective_size);
	if (rc == MC2_ERR_INPUT) {
		mc2::set_error("invalid nucleotide letter (reference: InvalidInputException, ChromosomeOneDigit.cpp:86-95)");
		return rc;
	}
	if (rc != MC2_OK) {
		mc2::set_error("mc2_encode_dna: sequence longer than 2^31-1");
		return rc;
	}
	*n_segs = segs.size();
	if (segs.size() > max_segs) {
		mc2::set_error("mc2_encode_dna: segs_out too small");
		return MC2_ERR_ARG;
	}
	for (size_t j = 0; j < segs.size(); j++) {
		segs_out[2 * j] = segs[j].s;
		segs_out[2 * j + 1] = segs[j].e;
	}
	return MC2_OK;
}

int mc2_encode_dna_batch(const char *text, const uint64_t *off, uint64_t n, char *codes_out, int32_t *segs_out,
			 uint64_t max_segs, uint64_t *seg_off_out, uint64_t *effective_sizes, int threads)
{
	if (!off || !seg_off_out || (n && (!text || !codes_out))) {
		mc2::set_error("mc2_encode_dna_batch: NULL argument");
		return MC2_ERR_ARG;
	}
	std::vector<std::vector<Seg>> all(n);
	int bad = 0;
	(void)threads;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads > 0 ? threads : 1)
	for (int64_t i = 0; i < (int64_t)n; i++) {
		uint64_t eff = 0;
		int rc = encode_one(text + off[i], off[i + 1] - off[i], codes_out + (off[i] - off[0]), all[i], &eff);
		if (effective_sizes) {
			effective_sizes[i] = eff;
		}
		if (rc != MC2_OK) {
#pragma omp atomic write
			bad = rc;
		}
	}
	if (bad) {
		mc2::set_error("mc2_encode_dna_batch: invalid nucleotide letter or over-long sequence");
		return bad;
	}
	uint64_t tot = 0;
	for (uint64_t i = 0; i < n; i++) {
		seg_off_out[i] = tot;
		tot += all[i].size();
	}
	seg_off_out[n] = tot;
	if (tot > max_segs) {
		mc2::set_error("mc2_encode_dna_batch: segs_out too small");
		return MC2_ERR_ARG;
	}
	for (uint64_t i = 0; i < n; i++) {
		for (size_t j = 0; j < all[i].size(); j++) {
			segs_out[2 * (seg_off_out[i] + j)] = all[i][j].s;
			segs_out[2 * (seg_off_out[i] + j) + 1] = all[i][j].e;
		}
	}
	return MC2_OK;
}

} // extern "C"
