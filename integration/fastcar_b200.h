// fastcar_b200.h -- the one call the relinked fastcar adds to src/fastcar/FC_Runner.cpp (integration/patch_fc_runner.py):
// work() offers its whole (query chunk x database chunk) block to the device before running the reference's double loop.
#pragma once
#include <fstream>
#include <string>
#include <vector>

#include "clutil/Point.h"
#include "predict/Predictor.h"

// work() of src/fastcar/FC_Runner.cpp:427-470 for one block: every query against the database points inside the length
// window, Predictor::close (classifier) and, for survivors, Predictor::similarity (regression), lines written to `out` in the
// reference's order.  Returns false when it declines (the reference's loop then runs): histogram types the device does not
// serve, or a predictor without a classifier.
template <class T>
bool mc2_batched_work(const std::vector<Point<T> *> &queries, const std::vector<Point<T> *> &pts, double similarity, Predictor<T> *pred,
		      const std::string &delim, std::ofstream &out, uintmax_t &num_pred_pos, bool format,
		      std::string (*format_header)(std::string));
