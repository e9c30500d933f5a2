#!/bin/bash
# One gpurun call of round 2: GPU tests, the bench line, the ncu launch list of the bench command and `ncu --set full`
# captures of the HBM kernels (each ncu command only after the same command exited 0 without ncu).
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh [tests] [bench] [launches] [full] [dense]'
# Everything lands in gpurun_out/; the summaries that are judged are copied to profiles/ by hand.
set -u
mkdir -p gpurun_out
what="${*:-tests bench launches full dense}"
TAG=${TAG:-r2}
SMALL="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-rowpart --no-nonsym --no-dense"
# gpurun brings back at most 64 MiB: a report is exported to CSV on the box (raw counters per launch, and the source page
# with per-line stall samples) and the .ncu-rep itself is dropped
export_rep() {
  [ -f $1.ncu-rep ] || return 0
  ncu -i $1.ncu-rep --page raw --csv > $1_raw.csv 2>/dev/null
  ncu -i $1.ncu-rep --page source --csv > $1_source.csv 2>/dev/null
  gzip -f $1_source.csv
  rm -f $1.ncu-rep
}
has() { case " $what " in *" $1 "*) return 0;; esac; return 1; }

if has lsap; then   # new cluster kernel: on its own, under a short limit, before anything else relies on the GPU
  timeout 180 python -m pytest tests/test_gpu_parity.py -x -q -k "lsap or hungarian" > gpurun_out/${TAG}_lsap.log 2>&1
  echo "lsap tests exit $?"; tail -5 gpurun_out/${TAG}_lsap.log
fi
if has tests; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputest.log 2>&1
  echo "gpu tests exit $?"; tail -3 gpurun_out/${TAG}_gputest.log
fi
if has bench; then
  timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  echo "bench exit $?"; cut -c1-600 gpurun_out/${TAG}_bench.json
fi
small_ok=1
if has launches || has full; then   # the plain run every ncu command below depends on (same command line, same files)
  timeout 300 $SMALL > gpurun_out/${TAG}_small.json 2> gpurun_out/${TAG}_small.err
  small_ok=$?; echo "plain small bench exit $small_ok"
fi
if has launches && [ $small_ok -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches_bench_p128.csv \
      $SMALL > gpurun_out/${TAG}_ncu_launches.log 2>&1
  echo "launch list exit $?"
fi
if has full && [ $small_ok -eq 0 ]; then
  for spec in "corr:k_filter_sell<.int.16, .int.4, .int.3,:40" "f32:k_filter_sell<.int.16, .int.4, .int.0,:40" "smooth:k_mean_filter<.int.3>:400"; do
    name=${spec%%:*}; rest=${spec#*:}; pat=${rest%:*}; skip=${rest##*:}
    timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$pat" -s $skip -c 3 \
        -f -o gpurun_out/${TAG}_full_$name $SMALL > gpurun_out/${TAG}_ncu_full_$name.log 2>&1
    echo "full $name exit $?"
    export_rep gpurun_out/${TAG}_full_$name
  done
fi
if has dense; then
  timeout 600 python tools/dense_evidence.py > gpurun_out/${TAG}_dense.json 2> gpurun_out/${TAG}_dense.err
  echo "dense exit $?"; cut -c1-400 gpurun_out/${TAG}_dense.json
  if has densencu; then
    timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:k_gram<.int.96|k_rotate<.int.96|k_pk_search<.int.3, .int.1" -c 8 \
        -f -o gpurun_out/${TAG}_full_dense python tools/dense_evidence.py --no-micro > gpurun_out/${TAG}_ncu_full_dense.log 2>&1
    echo "dense ncu exit $?"
    export_rep gpurun_out/${TAG}_full_dense
  fi
fi
if has ab; then
  timeout 600 python tools/kernel_ab.py > gpurun_out/${TAG}_kernel_ab.log 2>&1
  echo "kernel_ab exit $?"; tail -8 gpurun_out/${TAG}_kernel_ab.log
fi
if has knnncu; then
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_gram_wide|k_pk_search<.int.3" -c 10 \
      -f -o gpurun_out/${TAG}_full_gramknn python tools/dense_evidence.py --no-micro --k65-only > gpurun_out/${TAG}_ncu_full_gramknn.log 2>&1
  echo "gram/knn ncu exit $?"
  export_rep gpurun_out/${TAG}_full_gramknn
fi
if has latlist; then
  timeout 300 python tools/pair_latency.py --no-oracle > gpurun_out/${TAG}_pair_latency_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/${TAG}_launches_pair.csv \
      python tools/pair_latency.py --no-oracle > gpurun_out/${TAG}_ncu_pair.log 2>&1
  echo "pair launch list exit $?"
fi
if has latency; then
  timeout 300 python tools/pair_latency.py > gpurun_out/${TAG}_pair_latency.log 2>&1
  echo "latency exit $?"; cat gpurun_out/${TAG}_pair_latency.log
  timeout 300 python tools/pair_latency.py --fp64 --no-oracle > gpurun_out/${TAG}_pair_latency_fp64.log 2>&1
  echo "latency (fp64 passes) exit $?"; cat gpurun_out/${TAG}_pair_latency_fp64.log
fi
ls -la gpurun_out | tail -30
