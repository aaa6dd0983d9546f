#!/bin/bash
# Collect this round's measurement artefacts on the GPU box into gpurun_out/r02/ (summarised into profiles/ by
# tools/summarize_round.py).  Every ncu pass runs after the same command exited 0 without ncu.
set -u
O=${1:-gpurun_out/r02}; mkdir -p $O
NCU="ncu --set full --import-source on --clock-control none"
python tools/prof_step.py deepconn > $O/prof_step_deepconn.log 2>&1 || exit 1
python tools/prof_step.py narre > $O/prof_step_narre.log 2>&1 || exit 1
SPECS=("deepconn conv_tc2_kernel 4" "deepconn cmat_scatter_fast_kernel 4" "deepconn cmat_split_kernel 2" "deepconn cmat_table_gemm_kernel 2" \
       "deepconn cmat_weight_gemm_kernel 2" "deepconn clip_adam_kernel 2" "deepconn sumsq_kernel 2" "deepconn head_fwd2_kernel 2" \
       "deepconn head_bwd2_kernel 2" "deepconn gather_rows_v4_kernel 1" "deepconn conv_rowidx_kernel 4" "narre conv_doc_select_kernel 4" \
       "narre conv_tc2_kernel 4" "narre cmat_scatter_fast_kernel 4" "narre cmat_table_gemm_kernel 2" "narre cmat_weight_gemm_kernel 2" \
       "narre narre_attn_tc_fwd_kernel 2" "narre narre_attn_tc_bwd_kernel 2")
# "conv": only the kernels of the forward conv (the reports carry the source pages: ~10 MB each, and gpurun_out/ is capped at 64 MiB)
[ "${2:-full}" = conv ] && SPECS=("deepconn conv_tc2_kernel 4" "deepconn conv_rowidx_kernel 4" "narre conv_tc2_kernel 4")
for spec in "${SPECS[@]}"; do
  set -- $spec
  $NCU -k regex:$2 --launch-skip $3 -c 1 -f -o $O/${1}_$2 python tools/prof_step.py $1 6 > $O/ncu_${1}_$2.log 2>&1
done
for m in deepconn narre dual_att; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_${m}_step.csv \
      python bench.py --model $m --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-library-baseline --no-full-step --graphs off > $O/ncu_launch_$m.log 2>&1
done
ls -la $O | tail -40
