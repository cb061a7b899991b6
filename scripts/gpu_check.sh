#!/bin/bash
# One GPU lease = one call of this script (gpurun -- 'bash scripts/gpu_check.sh TAG [what ...]').
#   tests      all GPU tests (+ smoke)            quick     only the tests named in $EMRT_QUICK_TESTS
#   bench      default bench + reference arm      launches  ncu launch list of a 2-step bench -> TAG_launches.{csv,md}
#   gather     ncu --set full of the window gather ln        ncu --set full of the fused Linear + LayerNorm GEMM
#   train      cfg-4 training step                micro     shared-memory-pipe floor microbenchmark
#   multi      bench under torchrun at NGPU ranks (use with gpurun --gpus N)
# Everything lands in gpurun_out/TAG_*; the summaries worth keeping are copied to profiles/ by hand.
TAG=${1:-x}; shift
WHAT=${@:-tests bench launches}
mkdir -p gpurun_out
for w in $WHAT; do
case $w in
quick)
  timeout 900 python -m pytest $EMRT_QUICK_TESTS -m gpu -q -x --timeout 600 2>&1 | tail -25 | tee gpurun_out/${TAG}_quick.log ;;
tests)
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -25 > gpurun_out/${TAG}_pytest.log; tail -6 gpurun_out/${TAG}_pytest.log
  python __graft_entry__.py --smoke 2>&1 | tail -1 ;;
bench)
  timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; tail -c 200 gpurun_out/${TAG}_bench_reference.json ;;
launches)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
  python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.md; head -24 gpurun_out/${TAG}_launches.md ;;
gather)
  timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:msda_gather_fwd_win -s 1 -c 1 -o gpurun_out/${TAG}_gather_win \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1 ;;
ln)
  timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:linear_ln_tcgen05 -s 1 -c 1 -o gpurun_out/${TAG}_linear_ln \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1 ;;
op)
  # single-kernel capture: EMRT_OP = qproj | ffn1 | ffn2 | value, EMRT_OP_KERNEL = kernel-name regex (default: linear)
  timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:${EMRT_OP_KERNEL:-linear} -s 2 -c 1 -o gpurun_out/${TAG}_${EMRT_OP:-qproj} \
      python scripts/op_once.py ${EMRT_OP:-qproj} > /dev/null 2>&1; python scripts/op_once.py ${EMRT_OP:-qproj} ;;
train)
  timeout 600 python scripts/bench_train.py --steps 10 2>&1 | tail -1 | tee gpurun_out/${TAG}_train.json ;;
multi)
  # under `gpurun --gpus N`: the bench exactly as the driver launches it at N ranks (NGPU from the environment, default 2)
  N=${NGPU:-2}
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err; tail -c 2500 gpurun_out/${TAG}_bench_n${N}.json; tail -3 gpurun_out/${TAG}_bench_n${N}.err ;;
micro)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/lds_floor scripts/microbench/lds_gather_floor.cu && /tmp/lds_floor | tee gpurun_out/${TAG}_lds_floor.txt ;;
*) echo "unknown step $w" ;;
esac
done
