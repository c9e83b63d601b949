// CPU check of the host-side pieces of the device reader (integration/fasta_records.h) against the reference's own reader,
// compiled against the reference sources and linked with oracle/_ref/libmc2ref.so (TEST INFRASTRUCTURE).
//   test_fasta_records <file.fa>
// 1. split_fasta == ChromListMaker::makeChromOneDigitDnaList record for record (header, size);
// 2. encode_data_string over the reference's own segment list == the base string ChromosomeOneDigit::encode leaves
//    (or both reject the record's letters);
// 3. the effective size of `A...A + sequence` (doubled_for_find_k, segmented by the reference's Chromosome) == the effective
//    size of makeChromList's blank-prefilled Chromosome (what Runner::find_k averages).
#include <fstream>
#include <iostream>
#include <sstream>

#include "exception/InvalidInputException.h"
#include "nonltr/ChromListMaker.h"
#include "nonltr/Chromosome.h"
#include "nonltr/ChromosomeOneDigitDna.h"

#include "fasta_records.h"

using namespace nonltr;

int main(int argc, char **argv)
{
	if (argc < 2) {
		return 2;
	}
	std::ifstream in(argv[1], std::ios::binary);
	std::stringstream ss;
	ss << in.rdbuf();
	const std::string raw = ss.str();
	mc2i::FileRecords rec;
	if (!mc2i::split_fasta(raw, rec)) {
		std::cout << "DECLINED" << std::endl;
		return 0;
	}
	const uint64_t n = rec.headers.size();
	bool ref_threw = false;
	const std::vector<Chromosome *> *list = nullptr;
	ChromListMaker maker(argv[1], false);
	try {
		list = maker.makeChromOneDigitDnaList();
	} catch (InvalidInputException &e) {
		ref_threw = true;
	}
	bool ours_rejects = false;
	if (!ref_threw) {
		if (list->size() != n) {
			std::cout << "FAIL record count " << n << " vs " << list->size() << std::endl;
			return 1;
		}
	}
	for (uint64_t i = 0; i < n; i++) {
		std::string data(rec.text, rec.seq_off[i], rec.seq_off[i + 1] - rec.seq_off[i]);
		if (ref_threw) {
			// segment the record with the reference's plain Chromosome (no letter check) to get what encode would have seen
			std::string copy = data, hdr = rec.headers[i];
			Chromosome plain(copy, hdr);
			std::vector<int32_t> segs;
			for (const std::vector<int> *s : *plain.getSegment()) {
				segs.push_back(s->at(0));
				segs.push_back(s->at(1));
			}
			ours_rejects = ours_rejects || mc2i::encode_data_string(data, segs.data(), segs.size() / 2) != 0;
			continue;
		}
		Chromosome *c = list->at(i);
		if (c->getHeader() != rec.headers[i] || (uint64_t)c->size() != data.size()) {
			std::cout << "FAIL record " << i << " header/size" << std::endl;
			return 1;
		}
		std::vector<int32_t> segs;
		uint64_t eff = 0;
		for (const std::vector<int> *s : *c->getSegment()) {
			segs.push_back(s->at(0));
			segs.push_back(s->at(1));
			eff += (uint64_t)(s->at(1) - s->at(0) + 1);
		}
		if (eff != (uint64_t)c->getEffectiveSize()) {
			std::cout << "FAIL record " << i << " effective size" << std::endl;
			return 1;
		}
		const char bad = mc2i::encode_data_string(data, segs.data(), segs.size() / 2);
		if (bad || data != *c->getBase()) {
			std::cout << "FAIL record " << i << " data string" << std::endl;
			return 1;
		}
	}
	if (ref_threw) {
		std::cout << (ours_rejects ? "OK both reject" : "FAIL reference rejects, ours accepts") << std::endl;
		return ours_rejects ? 0 : 1;
	}
	// find_k
	ChromListMaker maker2(argv[1], false);
	const std::vector<Chromosome *> *plain = maker2.makeChromList();
	std::string doubled;
	std::vector<uint64_t> off;
	mc2i::doubled_for_find_k(rec, doubled, off);
	if (plain->size() != n) {
		std::cout << "FAIL makeChromList record count" << std::endl;
		return 1;
	}
	for (uint64_t i = 0; i < n; i++) {
		std::string d(doubled, off[i], off[i + 1] - off[i]), hdr = rec.headers[i];
		Chromosome c(d, hdr);
		if (c.getEffectiveSize() != plain->at(i)->getEffectiveSize()) {
			std::cout << "FAIL record " << i << " find_k effective size " << c.getEffectiveSize() << " vs "
				  << plain->at(i)->getEffectiveSize() << std::endl;
			return 1;
		}
	}
	std::cout << "OK " << n << " records" << std::endl;
	return 0;
}
