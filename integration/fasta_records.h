// fasta_records.h -- host-side pieces of the device reader (integration/GetPoints_b200.cpp), header-only so that the CPU
// test (tests/cpp/test_fasta_records.cpp) can hold them against the reference's own reader.
#ifndef MC2_FASTA_RECORDS_H
#define MC2_FASTA_RECORDS_H

#include <cctype>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace mc2i {

// One FASTA file split into records with the reader's line rules: lines end at \n, \r\n or \r (safe_getline,
// src/nonltr/ChromListMaker.cpp:24-47); a line starting with '>' opens a record and is its header, verbatim; lines starting
// with a blank or a tab are skipped; every other line is appended to the current record as it is (:131-158).
struct FileRecords {
	std::vector<std::string> headers;
	std::vector<uint64_t> seq_off; // [n + 1] offsets into text
	std::string text;              // the records' bases, concatenated, as written in the file
};

// false = a shape that trips the reader's own corner cases, left to the reference: bases before any header (it dereferences
// an unset pointer there), no record at all, or a record without bases (which shifts its size list, ChromListMaker.cpp:105-107)
inline bool split_fasta(const std::string &raw, FileRecords &rec)
{
	rec.headers.clear();
	rec.seq_off.assign(1, 0);
	rec.text.clear();
	rec.text.reserve(raw.size());
	const size_t sz = raw.size();
	size_t pos = 0;
	while (pos < sz) {
		size_t e = pos;
		while (e < sz && raw[e] != '\n' && raw[e] != '\r') {
			e++;
		}
		const char first = e > pos ? raw[pos] : '\0';
		if (first == '>') {
			if (!rec.headers.empty()) {
				rec.seq_off.push_back(rec.text.size());
			}
			rec.headers.push_back(raw.substr(pos, e - pos));
		} else if (first == ' ' || first == '\t') {
		} else if (e > pos) {
			if (rec.headers.empty()) {
				return false;
			}
			rec.text.append(raw, pos, e - pos);
		}
		pos = e;
		if (pos < sz) {
			pos += (raw[pos] == '\r' && pos + 1 < sz && raw[pos + 1] == '\n') ? 2 : 1;
		}
	}
	if (rec.headers.empty()) {
		return false;
	}
	rec.seq_off.push_back(rec.text.size());
	for (size_t i = 0; i + 1 < rec.seq_off.size(); i++) {
		if (rec.seq_off[i + 1] == rec.seq_off[i]) {
			return false;
		}
	}
	return true;
}

// the DNA code map of ChromosomeOneDigitDna::buildCodes (src/nonltr/ChromosomeOneDigitDna.cpp:48-68); -1 = not a nucleotide
struct CodeTable {
	signed char code[256];
	CodeTable()
	{
		std::memset(code, -1, sizeof code);
		const char *letters = "ACGTRYMKSWHBVDNX";
		const signed char val[] = {0, 1, 2, 3, 2, 1, 0, 3, 2, 3, 1, 3, 0, 3, 1, 2};
		for (int i = 0; letters[i]; i++) {
			code[(unsigned char)letters[i]] = val[i];
		}
	}
};

// What ChromosomeOneDigit::encode leaves in `base` (src/nonltr/ChromosomeOneDigit.cpp:79-133) after Chromosome::help has
// upper-cased it: with at least one segment every letter becomes its code except an N outside the segments, which stays 'N';
// without segments the upper-cased letters stay as they are.  segs = n_segs inclusive, sorted [start, end] pairs.
// Returns 0, or the first letter that is not a nucleotide (the reference throws InvalidInputException there).
inline char encode_data_string(std::string &data, const int32_t *segs, uint64_t n_segs)
{
	static const CodeTable table;
	for (char &c : data) {
		c = (char)toupper((unsigned char)c);
	}
	if (n_segs == 0) {
		return 0;
	}
	uint64_t sg = 0;
	for (size_t j = 0; j < data.size(); j++) {
		while (sg < n_segs && (int64_t)j > segs[2 * sg + 1]) {
			sg++;
		}
		const bool inside = sg < n_segs && (int64_t)j >= segs[2 * sg];
		const char c = data[j];
		if (!inside && c == 'N') {
			continue;
		}
		const signed char code = table.code[(unsigned char)c];
		if (code < 0) {
			return c;
		}
		data[j] = (char)code;
	}
	return 0;
}

// Runner::find_k measures Chromosome objects that makeChromList pre-fills with `size` blanks and then APPENDS the sequence to
// (src/nonltr/ChromListMaker.cpp:72,87 vs src/nonltr/Chromosome.cpp:18-25, 88-97).  Their segmentation only
// distinguishes N from not-N, so `A...A + sequence` with every non-N letter folded to A segments the same way -- and is valid
// input for the letter-checking device contract.
inline void doubled_for_find_k(const FileRecords &rec, std::string &doubled, std::vector<uint64_t> &off)
{
	const uint64_t n = rec.headers.size();
	doubled.clear();
	doubled.reserve(2 * rec.text.size());
	off.assign(n + 1, 0);
	for (uint64_t i = 0; i < n; i++) {
		const uint64_t len = rec.seq_off[i + 1] - rec.seq_off[i];
		doubled.append(len, 'A');
		doubled.append(rec.text, rec.seq_off[i], len);
		off[i + 1] = doubled.size();
	}
	for (char &c : doubled) {
		c = (c == 'N' || c == 'n') ? 'N' : 'A';
	}
}

} // namespace mc2i

#endif
