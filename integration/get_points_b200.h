// get_points_b200.h -- the entry point the edited src/cluster/CRunner.cpp calls (integration/patch_crunner.py):
// the k-mer histograms of every sequence of one FASTA file in one device batch instead of one Loader<T>::get_point per
// sequence.  Returns false when it declines (MC2_NO_BATCH=1, a histogram type the device path does not serve, or a file
// below MC2_K1_MIN_BASES bases -- default 32 Mi -- where the host loop ends before the CUDA context is up); the caller then
// runs the reference's own loop.  Defined in integration/GetPoints_b200.cpp.
#ifndef MC2_GET_POINTS_B200_H
#define MC2_GET_POINTS_B200_H

#include <cstdint>
#include <string>
#include <vector>

#include "clutil/Point.h"
#include "nonltr/Chromosome.h"

// for elt in chroms: points.push_back(Loader<T>::get_point(elt, id, k))      (src/cluster/CRunner.cpp:523-528)
template <class T>
bool mc2_batched_get_points(const std::vector<Chromosome *> &chroms, uintmax_t &id, int k, std::vector<Point<T> *> &points);


// The whole per-file body of get_points (src/cluster/CRunner.cpp:521-528: ChromListMaker + makeChromOneDigitDnaList + one
// get_point per sequence) for a multi-record FASTA: the file is split into records on the host (the reader's line rules,
// src/nonltr/ChromListMaker.cpp:24-47, 117-165), and the input contract itself -- upper-casing, N-run segmentation, bridging,
// dropping, 1 Mbp pieces, letter coding, ChromosomeOneDigit::finalize -- runs on the device (mc2_seqs_from_text) in front of
// K1.  Returns false when it declines: MC2_NO_BATCH=1, --single-file, a file below MC2_K1_MIN_BASES bytes, or a file whose
// shape trips the reader's own corner cases (no leading header, a record without bases).
// Runner::find_k's per-file body (src/cluster/CRunner.cpp:484-493) and the per-file body of Runner::run's histogram-width
// detection (:61-81), from the same split records: the reference reads every input file three times.
bool mc2_batched_effective_length(const std::string &fasta, bool is_single_file, unsigned long long &sum_effective,
				  unsigned long long &n_records);
bool mc2_batched_largest_count(const std::string &fasta, bool is_single_file, int k, uint64_t &largest);

template <class T>
bool mc2_batched_read_points(const std::string &fasta, bool is_single_file, uintmax_t &id, int k, std::vector<Point<T> *> &points);

#endif
