#!/bin/bash
# Build an alternative libkmers_b200.so with extra -D flags:  bash scripts/build_variant.sh NAME -DKMB_X=1 ...   -> build/exp/NAME.so
# (load it with KMERS_B200_SO=build/exp/NAME.so; kernel experiments only)
name=$1; shift
out=build/exp/$name; mkdir -p $out
cd kmers_b200/csrc
objs=""
for f in *.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c -o ../../$out/${f%.cu}.o $f &
  objs="$objs $out/${f%.cu}.o"
done
for f in *.cpp; do
  g++ -O3 -std=c++17 -fPIC -pthread -fvisibility=hidden -c -o ../../$out/${f%.cpp}.o $f &
  objs="$objs $out/${f%.cpp}.o"
done
wait
cd ../..
nvcc -shared -Wno-deprecated-gpu-targets -o build/exp/$name.so $objs -ldl -lpthread && rm -rf $out && ls -la build/exp/$name.so
