#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of one proof + full ncu captures of the dominant kernels.
# Reports are exported to CSV on the box (the .ncu-rep files embed the whole cubin and exceed the gpurun_out quota).
#   tools/capture_profiles.sh <tag> [log_n]
tag=${1:-r02}; logn=${2:-20}
out=gpurun_out; tmp=/tmp/ncu_$tag; mkdir -p $out $tmp
# launch list of the bench command itself (B200_PROFILING.md): every launch with its device time
BENCH_NO_SMI=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_prove.log 2>&1
capture() {  # name regex skip count
    ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o $tmp/$1 -f python tools/prove_once.py $logn 2 >> $out/${tag}_prove.log 2>&1
    ncu -i $tmp/$1.ncu-rep --page raw --csv > $out/${tag}_$1_raw.csv 2>/dev/null
    ncu -i $tmp/$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $out/${tag}_$1_source.csv.gz
    ncu -i $tmp/$1.ncu-rep --page details 2>/dev/null | head -400 > $out/${tag}_$1_details.txt
}
# one proof launches 16 ntt_pass_kernel (2 interpolation + 4 x 2 trace LDE + 2 + 2 + 2): skip the first proof and the interpolation of the
# second, capture one (strided pass, contiguous pass) pair of the trace LDE
capture ntt ntt_pass_kernel 18 2
capture air constraint_eval 1 1
capture hash "^hash_rows_kernel" 1 1
capture merkle merkle_level 20 1
du -sh $out
