#!/bin/bash
# First multi-GPU call of the NEXT session: everything written after the round-1 GPU budget was spent.
# Ring (periodic box on N GPUs): correctness against the ring of one slab and the oracle,
# then config 4's weak-scaling point at N GPUs. usage (inside gpurun --gpus N): scripts/gpu_ring_n.sh N [n_side_bench]
N=${1:-2}; NS=${2:-256}; OUT=gpurun_out/ring_n$N; mkdir -p $OUT
export SPHB200_CHECK_EXCHANGE=1   # first runs of the ring over NCCL: a plane-size mismatch stops with a message instead of a hang
# the two single-GPU ring tests that have not been run yet
SPHB200_RUN_UNVERIFIED=1 timeout 200 python -m pytest tests/test_gpu_zz_periodic_ring.py -m gpu -q > $OUT/pytest_unverified.log 2>&1; tail -3 $OUT/pytest_unverified.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $RUN --master-port 29533 tests/multi_gpu_check_ring.py --n-side 32 --outer 12 --drift 2.0 --out $OUT/ring_check.json \
    > $OUT/ring_check.log 2>&1; echo "ring check rc=$?"
grep RING_CHECK $OUT/ring_check.log | cut -c1-1500
SPHB200_CHECK_EXCHANGE=0 timeout 300 $RUN --master-port 29544 scripts/config4_bench.py $NS 6 > $OUT/config4.log 2>&1; echo "config4 rc=$?"
tail -3 $OUT/config4.log
# the complete reference case file on N GPUs (correction variants + free-surface indication + pressure probes): first run
SPHB200_CHECK_EXCHANGE=1 timeout 240 $RUN --master-port 29555 tests/multi_gpu_check.py --dp 0.05 --outer 20 --recut-interval 7 \
    --correction --surface-indicator --observers --out $OUT/complete_case_check.json > $OUT/complete_case_check.log 2>&1; echo "complete case rc=$?"
grep MULTI_GPU_CHECK $OUT/complete_case_check.log | cut -c1-900; tail -2 $OUT/complete_case_check.log | cut -c1-300
