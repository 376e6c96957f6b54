#!/bin/bash
# tools/sanitize.sh -- compute-sanitizer over the small-configuration GPU parity tests.
# memcheck on every suite, racecheck on the kernels with shared-memory queues / rings (both neighbour kernels, the
# fused neighbour+CNA kernel), union-find and atomicMin claims.  Summaries -> gpurun_out/sanitizer_*.txt
set -u
OUT=gpurun_out
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool tests...
    local name=$1 tool=$2; shift 2
    timeout 1500 $CS --tool $tool --print-limit 20 --error-exitcode 0 --log-file $OUT/sanitizer_${name}_${tool}.log \
        python -m pytest "$@" -x -q -p no:cacheprovider > $OUT/sanitizer_${name}_${tool}.pytest.txt 2>&1
    echo "== $name $tool: pytest: $(tail -1 $OUT/sanitizer_${name}_${tool}.pytest.txt)" >> $OUT/sanitizer_summary.txt
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/sanitizer_${name}_${tool}.log | tail -1 >> $OUT/sanitizer_summary.txt
}
: > $OUT/sanitizer_summary.txt
run neighbor memcheck tests/test_gpu_neighbor.py tests/test_gpu_knn.py tests/test_gpu_fused.py
run descriptors memcheck tests/test_gpu_descriptors.py tests/test_gpu_list_consumers.py tests/test_gpu_ids.py tests/test_gpu_chill_bond.py
run system memcheck tests/test_gpu_system.py tests/test_gpu_slab.py tests/test_gpu_distributed.py tests/test_gpu_builders.py
run ptm memcheck tests/test_gpu_ptm.py tests/test_gpu_planar_faults.py
run group_voronoi memcheck tests/test_gpu_group.py tests/test_gpu_voronoi.py
run neighbor racecheck tests/test_gpu_neighbor.py tests/test_gpu_fused.py
MDB_NEIGHBOR=tiled_v1 run neighbor_v1 racecheck tests/test_gpu_neighbor.py
MDB_NEIGHBOR=coop run neighbor_coop racecheck tests/test_gpu_neighbor.py
run group racecheck tests/test_gpu_group.py -k "labels_equal_reference and (fcc_hot or shuffled or gas)"
run sort_staged racecheck tests/test_gpu_neighbor.py tests/test_gpu_descriptors.py -k "sort_and_descriptors or host_pointer_dropins"
run cluster_ids racecheck tests/test_gpu_ids.py tests/test_gpu_list_consumers.py -k "cluster or ids or diamond"
cat $OUT/sanitizer_summary.txt
