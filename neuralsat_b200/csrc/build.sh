#!/bin/bash
# Builds libcrown_b200.so (sm_100a only) in-tree.  Usage: build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall $ARCH $*"
OUT=../libcrown_b200.so
SRCS="crown_kernels.cu crown_api.cu"
SRCS="$SRCS crown_tc.cu"
[ -f crown_chain.cu ] && SRCS="$SRCS crown_chain.cu"
SRCS="$SRCS crown_sshape.cu crown_chain_grad.cu crown_conv.cu crown_conv_tc.cu crown_store.cu"
OBJS=""
PIDS=""
for f in $SRCS; do
  [ -f "$f" ] || continue
  o="${f%.cu}.o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ crown_kernels.cuh -nt "$o" ] || [ crown_tc_common.cuh -nt "$o" ] || [ crown_chain_common.cuh -nt "$o" ] || [ ../../include/crown_b200.h -nt "$o" ]; then
    rm -f "$o"            # a stale object must never satisfy the link step after a failed compile
    $NVCC $FLAGS -c "$f" -o "$o" &
    PIDS="$PIDS $!"
  fi
  OBJS="$OBJS $o"
done
for pid in $PIDS; do wait $pid || { echo "compile failed (pid $pid)"; exit 1; }; done
for o in $OBJS; do [ -f "$o" ] || { echo "compile failed: $o"; exit 1; }; done
$NVCC -shared $ARCH -o $OUT $OBJS -lcudart
echo "built $(realpath $OUT)"
