#!/bin/bash
# compute-sanitizer over the kernels added or restructured in round 2 (tile pair / virial kernels with their shared-memory
# partner slots and folds, packed-table decode + verify kernel, CUDA-graph replay, zero-copy transpose, the NCCL entry
# points with one rank -- incl. the pipelined exchange --, the rho kernel reading the page-locked beads array in the
# graph, the persistent tau-correlation kernel).  Small tests only: the tools slow kernels down 10-100x.
TAG=${1:-r02s}
OUT=gpurun_out
mkdir -p $OUT
SEL='test_pair_sums_ragged_shapes or test_pair_table_index_is_exact or test_pair_table_edges or (test_pair_sums and C1) or test_fused_single_walker_call_graph_replay or test_pinned_source_takes_device_transpose_path or test_graph_replay_leaves_the_transposed_beads or (test_tau_correlation_kernels and (254 or 126 or 62))'
SEL2='test_virial_sums_vs_oracle or test_force_kernels_with_several_particles_per_thread'
for tool in memcheck racecheck initcheck synccheck; do
    echo "== $tool" | tee -a $OUT/${TAG}_sanitizer.txt
    timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" 2>&1 | tail -6 | tee -a $OUT/${TAG}_sanitizer.txt
    timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_variants.py tests/test_comm.py -x -q -m gpu -k "$SEL2 or single_rank" 2>&1 | tail -6 | tee -a $OUT/${TAG}_sanitizer.txt
done
