// update_batch_b200.h -- the two entry points the edited src/cluster/ClusterFactory.cpp calls (see
// integration/patch_cluster_factory.py and INTEGRATION.md): one device call per pass of the update stage instead of one per
// center.  Both return false when they decline (MC2_NO_BATCH=1, or no device mirror yet); the caller then runs the
// reference's own per-center loop.  Defined in integration/Trainer_b200.cpp.
#ifndef MC2_UPDATE_BATCH_B200_H
#define MC2_UPDATE_BATCH_B200_H

#include <vector>

#include "cluster/Center.h"
#include "cluster/Trainer.h"

// for j in [0, part.size()): mean_shift_update(part, j, trn, delta)   (src/cluster/ClusterFactory.cpp:288-335, 639-642, 648-651)
template <class T>
bool mc2_batched_update(std::vector<Center<T>> &part, const Trainer<T> &trn, int delta);

// merge(part, trn, delta, bandwidth)                                   (src/cluster/ClusterFactory.cpp:382-401, 643)
template <class T>
bool mc2_batched_merge(std::vector<Center<T>> &part, const Trainer<T> &trn, int delta);

#endif
