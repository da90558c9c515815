#!/bin/bash
# Build variants of libtrgt_b200.so that differ by -D flags (here, no GPU needed) into build/ab/ (which travels to the box; gpurun_out/ does not), and -- under
# gpurun -- time each one's resident pass.  Usage:
#   tools/ab_variants.sh build name1 "-DX=1" name2 "-DX=2" ...     (in the container)
#   tools/ab_variants.sh run name1 name2 ...                        (on the GPU box; prints kernel times)
mode=$1; shift; mkdir -p build/ab
cd "$(dirname "$0")/.."
if [ "$mode" = build ]; then
  while [ $# -gt 0 ]; do
    name=$1; flags=$2; shift 2
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall -shared $flags \
      -o build/ab/$name.so trgt_b200/csrc/engine.cu || exit 1
    echo built $name
  done
else
  mkdir -p gpurun_out/ab build/ab; cp trgt_b200/libtrgt_b200.so build/ab/_orig.so
  for name in "$@"; do
    cp build/ab/$name.so trgt_b200/libtrgt_b200.so
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --resident-only > gpurun_out/ab/$name.json 2> gpurun_out/ab/$name.err
    echo "== $name"; python tools/show_bench.py gpurun_out/ab/$name.json | grep -E "value|k_flank|k_e2e|k_hmm|k_wfa_score_warp" | cut -c1-60
  done
  cp build/ab/_orig.so trgt_b200/libtrgt_b200.so
fi
