// test_shim.cpp — exercises meshclust2_b200/host/mc2_shim.hpp (the reference-named C++ classes over the C ABI) against a
// fixture file written by tests/test_shim.py from the reference-generated golden vectors.
//   usage: test_shim <fixture.txt> <weights.txt>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include "../../meshclust2_b200/host/mc2_shim.hpp"

using namespace mc2shim;

static int g_fail = 0, g_checks = 0;
#define CHECK(cond, msg)                                                     \
	do {                                                                 \
		g_checks++;                                                  \
		if (!(cond)) {                                               \
			g_fail++;                                            \
			std::cerr << "FAIL " << msg << " (" #cond ")\n";     \
		}                                                            \
	} while (0)

static bool rel_close(double a, double b, double tol)
{
	if (a == b) return true;
	return std::fabs(a - b) <= tol * std::fmax(std::fabs(b), 1e-300) || std::fabs(a - b) <= tol;
}

int main(int argc, char **argv)
{
	if (argc < 3) {
		std::cerr << "usage: test_shim fixture weights\n";
		return 2;
	}
	std::ifstream in(argv[1]);
	std::string tag;
	size_t n, N, m, S;
	int k;
	in >> tag >> n >> k >> N;
	std::vector<std::string> seqs(n), headers(n);
	for (size_t i = 0; i < n; i++) {
		in >> seqs[i];
		headers[i] = ">s" + std::to_string(i);
	}
	std::vector<std::vector<int>> hist(n, std::vector<int>(N));
	std::vector<uint64_t> len(n), mag(n);
	in >> tag;
	for (size_t i = 0; i < n; i++) {
		in >> len[i] >> mag[i];
		for (size_t b = 0; b < N; b++) in >> hist[i][b];
	}
	in >> tag >> m >> S;
	std::vector<uint64_t> ia(m), ib(m);
	std::vector<double> score(m);
	std::vector<int> close(m);
	std::vector<std::vector<double>> cache(m, std::vector<double>(S));
	for (size_t j = 0; j < m; j++) {
		in >> ia[j] >> ib[j] >> score[j] >> close[j];
		for (size_t s = 0; s < S; s++) in >> cache[j][s];
	}
	size_t ng;
	in >> tag >> ng;
	std::vector<uint64_t> gq(ng);
	std::vector<long> gbest(ng);
	std::vector<int> gmin(ng);
	size_t ncand;
	in >> ncand;
	std::vector<uint64_t> cand(ncand);
	for (auto &c : cand) in >> c;
	std::vector<std::vector<int>> gmarks(ng, std::vector<int>(ncand));
	for (size_t t = 0; t < ng; t++) {
		in >> gq[t] >> gbest[t] >> gmin[t];
		for (auto &x : gmarks[t]) in >> x;
	}
	if (!in) {
		std::cerr << "bad fixture\n";
		return 2;
	}

	// ---- Loader<T>::get_points (+ device set) ----
	uintmax_t id = 0;
	std::unique_ptr<PointSet<uint8_t>> dev;
	auto pts = Loader<uint8_t>::get_points(headers, seqs, id, k, &dev, 2);
	CHECK(id == n, "ids advance by the number of points");
	for (size_t i = 0; i < n; i++) {
		bool same = true;
		for (size_t b = 0; b < N; b++) same = same && (int)pts[i]->points[b] == hist[i][b];
		CHECK(same, "histogram " << i);
		CHECK(pts[i]->get_length() == len[i] && pts[i]->getPseudoMagnitude() == mag[i], "length/mag " << i);
		CHECK(pts[i]->get_id() == i && pts[i]->getK() == k, "id/k " << i);
	}
	// single get_point with the ACGT-stripping overload
	{
		uintmax_t id2 = 7;
		auto *p = Loader<uint8_t>::get_point(">x", "ACGTNNNNACGTACGTxxACGTACGTACGTACGTACGT", id2, 3);
		CHECK(p->get_length() == 32 && id2 == 8 && p->get_id() == 7, "get_point(header, string) strips non-ACGT");
		delete p;
	}
	// ---- width detection (Runner::run) ----
	{
		int eb = 0;
		uint64_t lc = Loader<uint8_t>::largest_count(seqs, k, &eb);
		unsigned want = 0;
		for (size_t i = 0; i < n; i++) {
			for (size_t b = 0; b < N; b++) want = std::max<unsigned>(want, (unsigned)hist[i][b]);
		}
		CHECK(lc == want && eb == 1, "largest count " << lc << " vs " << want); // no saturation in the fixture: max bin = 1 + multiplicity
		std::vector<std::string> rep{std::string(400, 'A'), seqs[0]};
		lc = Loader<uint8_t>::largest_count(rep, k, &eb);
		CHECK(lc == (uint64_t)(400 - k + 1 + 1) && eb == 2, "poly-A needs 16 bits: " << lc);
	}
	// ---- KmerHashTable ----
	{
		std::string s = seqs[0].substr(0, 200);
		std::vector<char> codes(s.size());
		std::vector<int32_t> sg(2 * 8);
		uint64_t nseg = 0, eff = 0;
		check(mc2_encode_dna(s.data(), s.size(), codes.data(), sg.data(), 8, &nseg, &eff));
		KmerHashTable<unsigned long, uint8_t> t(3, 1);
		int r1 = t.wholesaleIncrementNoOverflow(codes.data(), 0, 100);
		int r2 = t.wholesaleIncrementNoOverflow(codes.data(), 101, 197);
		std::vector<unsigned> want(64, 1);
		for (int p = 0; p <= 197; p++) {
			int h = codes[p] * 16 + codes[p + 1] * 4 + codes[p + 2];
			if (want[h] < 255) want[h]++;
		}
		bool same = true;
		for (int h = 0; h < 64; h++) same = same && t.getValues()[h] == want[h];
		CHECK(same && r1 == 0 && r2 == 0 && t.getMaxTableSize() == 64 && t.getK() == 3, "KmerHashTable cumulative increments");
		KmerHashTable<unsigned long, uint8_t> sat(1, 250);
		std::vector<char> aaaa(64, 0);
		CHECK(sat.wholesaleIncrementNoOverflow(aaaa.data(), 0, 63) == -1 && sat.getValues()[0] == 255 && sat.getValues()[1] == 250,
		      "KmerHashTable saturation returns -1");
	}
	// ---- Predictor<T>(file): close / batch / get_close ----
	Predictor<uint8_t> pred(argv[2]);
	CHECK(pred.get_k() == k && pred.get_id() == 0.9 && pred.get_datatype() == "uint8_t" && pred.get_mode() == 1, "weights header");
	{
		std::vector<DivergencePoint<uint8_t> *> a, b;
		for (size_t j = 0; j < m; j++) {
			a.push_back(pts[ia[j]]);
			b.push_back(pts[ib[j]]);
		}
		auto c = pred.close_batch(a, b);
		size_t bad = 0;
		for (size_t j = 0; j < m; j++) bad += (int)c[j] != close[j] && std::fabs(score[j] - 0.5) > 1e-9;
		CHECK(bad == 0, "close_batch flags");
		CHECK(pred.close(pts[ia[0]], pts[ib[0]]) == (close[0] != 0), "close single");
		bool threw = false;
		try {
			pred.similarity(pts[0], pts[1]);
		} catch (const char *) {
			threw = true;
		}
		CHECK(threw, "similarity without a regression model throws like the reference");
	}
	for (size_t t = 0; t < ng; t++) {
		std::vector<uint8_t> marks;
		auto r = pred.get_close(*dev, gq[t], cand, marks);
		bool same = true;
		for (size_t j = 0; j < ncand; j++) same = same && (int)marks[j] == gmarks[t][j];
		CHECK(std::get<0>(r) == gbest[t] && std::get<2>(r) == (gmin[t] != 0) && same, "get_close " << t);
	}
	// ---- Feature<T>: add_feature / set_normal / compute / operator() with the model's own singles ----
	{
		const mc2_model_desc &d = pred.get_class_desc();
		Feature<uint8_t> feat(k);
		for (int c = 0; c < d.n_combos; c++) {
			uint64_t fl = 0;
			for (int t = 0; t < d.combo_nidx[c]; t++) fl |= d.single_flag[d.combo_idx[c][t]];
			Combo cb = d.combo_kind[c] == 0 ? Combo::xy : d.combo_kind[c] == 1 ? Combo::xy2 : d.combo_kind[c] == 2 ? Combo::x2y : Combo::x2y2;
			feat.add_feature(fl, cb);
		}
		for (int s = 0; s < d.n_singles; s++) feat.set_normal(d.single_flag[s], d.single_min[s], d.single_max[s]);
		feat.finalize();
		CHECK(feat.size() == (size_t)d.n_combos && feat.get_lookup().size() == S, "Feature shape");
		for (size_t j = 0; j < 8; j++) {
			auto cch = feat.compute(*pts[ia[j]], *pts[ib[j]]);
			bool ok = true;
			for (size_t s = 0; s < S; s++) ok = ok && rel_close(cch[s], cache[j][s], 1e-9);
			double sum = d.weight[0];
			for (int c = 0; c < d.n_combos; c++) sum += d.weight[c + 1] * feat(c, cch);
			CHECK(ok && std::fabs(Predictor<uint8_t>::classify_sum(sum) - score[j]) <= 1e-9, "Feature::compute + combos + classify_sum " << j);
		}
		// normalize(): min/max over a batch of raw singles, as Feature::normalize does
		Feature<uint8_t> f2(k);
		f2.add_feature(MC2_FEAT_MANHATTAN | MC2_FEAT_EMD, Combo::xy);
		std::vector<pra<uint8_t>> prs;
		for (size_t j = 0; j < 16; j++) prs.push_back(pra<uint8_t>{pts[ia[j]], pts[ib[j]], 0});
		f2.normalize(prs);
		auto raw = f2.raw_batch(prs);
		double mn = 1e300, mx = -1e300;
		for (size_t j = 0; j < 16; j++) {
			mn = std::fmin(mn, raw[2 * j]);
			mx = std::fmax(mx, raw[2 * j]);
		}
		CHECK(f2.get_normal(MC2_FEAT_MANHATTAN).first == mn && f2.get_normal(MC2_FEAT_MANHATTAN).second == mx, "Feature::normalize");
	}
	// ---- DivergencePoint clone / set semantics (stale mag, quirk Q4) ----
	{
		auto *c = pts[2]->clone();
		uint64_t m2 = c->getPseudoMagnitude();
		c->set(*pts[5]);
		CHECK(c->getPseudoMagnitude() == m2 && c->points == pts[5]->points && c->get_length() == pts[5]->get_length(),
		      "set() keeps the stale magnitude");
		delete c;
	}
	for (auto p : pts) delete p;
	std::cout << (g_fail ? "FAILED " : "OK ") << g_checks << " checks, " << g_fail << " failures\n";
	return g_fail ? 1 : 0;
}
