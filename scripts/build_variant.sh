#!/bin/bash
# build_variant.sh NAME "<extra nvcc defines>": tuning helper -- compiles csrc/solve.cu (or csrc/$SRC.cu, e.g. SRC=solve_ic) with extra -D flags and
# links it with the regular objects into optical-flow-python_b200/variants/lib_NAME.so (select with B200FLOW_LIB=...)
set -e
cd "$(dirname "$0")/../optical-flow-python_b200"
python build.py > /dev/null
mkdir -p variants
SRC=${SRC:-solve}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC --expt-relaxed-constexpr \
  $2 -c csrc/$SRC.cu -o variants/${SRC}_$1.o
OBJS=""
for o in pre warp solve solve_ic filter eval pipeline api; do
  if [ "$o" == "$SRC" ]; then OBJS="$OBJS variants/${SRC}_$1.o"; else OBJS="$OBJS build/$o.o"; fi
done
nvcc -shared -o variants/lib_$1.so $OBJS -gencode arch=compute_100a,code=sm_100a
echo variants/lib_$1.so
