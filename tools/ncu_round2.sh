#!/usr/bin/env bash
# ncu captures of the HBM-bound passes and hot kernels the round-1 review asked for (run on the GPU box).
# One invocation per kernel family, a few launches each; the raw pages are exported to CSV on the box and the
# reports deleted (gpurun_out/ is capped at 64 MiB) except the attention one.
set -uo pipefail
OUT=gpurun_out/ncu_r2; mkdir -p "$OUT"
LIGHT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy"
cap() {  # name target regex count mode
  local mode="--set full --import-source on"; [ "$5" = light ] && mode="$LIGHT"
  timeout 600 ncu --profile-from-start off $mode --clock-control none -k "regex:$3" -c "$4" \
    -f -o "$OUT/$1" python tools/ncu_step_target.py "$2" > "$OUT/$1.log" 2>&1 || tail -3 "$OUT/$1.log"
  ncu -i "$OUT/$1.ncu-rep" --page raw --csv > "$OUT/$1.raw.csv" 2>/dev/null
  [ "$1" = dit_attn ] || rm -f "$OUT/$1.ncu-rep"
}
cap dit_norm  dit 'ln_affine|scale_rows|unpatchify' 4 light
cap dit_attn  dit 'attn_persist|attn_combine' 3 full
cap dit_gemm  dit 'gemm_tc_kernel' 9 light
cap vae_norm  vae 'vae_norm|vae_upsample|vae_cast' 6 light
cap vae_conv  vae 'gemm_tc_kernel' 10 light
cap bwd_pass  bwd 'attn_ds_tile|attn_rowstat|colsum_stage1|transpose_h|ln_bwd|rms_rope_bwd' 10 light
ls -la "$OUT"
