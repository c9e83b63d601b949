#!/bin/bash
# compute-sanitizer passes over the CUDA path (run on a GPU box from the repo root); rounds 1 and 2: 0 errors / 0 hazards.
set -e
compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()"
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_pairs.py tests/test_gpu_ingest.py tests/test_gpu_k3.py \
    tests/test_gpu_count.py -m gpu -x -q -k "not full_size and not cfg4"
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_update_batch.py -m gpu -x -q
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_tile_sweep.py tests/test_gpu_scan_server.py -m gpu -x -q \
    -k "same_set_many or trained_models or scan or wide_rows_vs_oracle or declines"
compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_count.py tests/test_gpu_ingest.py -m gpu -x -q \
    -k "golden_histograms or auto_width or k8_packed or segments_equal"
