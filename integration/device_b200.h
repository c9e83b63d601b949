// device_b200.h -- the one CUDA context the relinked meshclust2 uses (integration/Trainer_b200.cpp for the candidate scans
// and the update stage, integration/GetPoints_b200.cpp for the k-mer histograms), and the lock that serialises its callers.
#ifndef MC2_DEVICE_B200_H
#define MC2_DEVICE_B200_H

#include <mutex>

#include "meshclust2_b200.h"

namespace mc2i {

// Creating the CUDA context (driver initialisation, loading the library's kernels, the page-locked staging areas) takes about
// a second and depends on nothing the program computes, so a thread started before main() does it while the FASTA file is
// read; shared_ctx() joins that thread (MC2_NO_PREWARM=1: created on first use instead).  Throws std::runtime_error when no
// sm_100 device is available -- there is no CPU fallback.  The context lives until the process ends.
mc2_ctx *shared_ctx();

// one GPU context, one caller at a time (get_points runs under `omp parallel for` over files, the update stage under one
// over centers)
std::mutex &device_mutex();

// MC2_NO_BATCH=1: the batched entry points decline and the reference's own loops run (through the per-call boundary)
bool batching_enabled();

void ok(int rc); // throws std::runtime_error(mc2_last_error()) unless rc == MC2_OK

} // namespace mc2i

#endif
