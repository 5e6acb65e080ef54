#!/usr/bin/env bash
# Builds libb200dit.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libb200dit.so"
OBJ="${HERE}/../build"
mkdir -p "${OBJ}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden)
SRCS=(gemm_tc gemm_tc_pair conv_tc attn_tc attn_bwd_tc elementwise backward_kernels profile dit_engine dit_backward vae_engine vae_kernels disc_engine capi)
pids=()
for s in "${SRCS[@]}"; do
  [ -f "${HERE}/${s}.cu" ] || continue
  if [ ! -f "${OBJ}/${s}.o" ] || [ -n "$(find "${HERE}" -newer "${OBJ}/${s}.o" \( -name '*.cu' -o -name '*.cuh' -o -name '*.h' \) -print -quit)" ] \
     || [ "${HERE}/../../include/b200dit.h" -nt "${OBJ}/${s}.o" ]; then
    "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${s}.cu" -o "${OBJ}/${s}.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
objs=()
for s in "${SRCS[@]}"; do [ -f "${OBJ}/${s}.o" ] && objs+=("${OBJ}/${s}.o"); done
"${NVCC}" -shared -o "${OUT}" "${objs[@]}" -Xlinker --exclude-libs=ALL
echo "built ${OUT}"
