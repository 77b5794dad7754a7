#!/bin/bash
# build_variant.sh NAME "-DFOO -DBAR=1": compiles csrc/attn_tc.cu with the extra defines and links a complete library
# orbit2_b200/libo2b200_NAME.so from it plus the regular objects (A/B kernel experiments: run with O2B200_LIB=that path).
set -e
cd "$(dirname "$0")/../orbit2_b200"
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $@ -c csrc/attn_tc.cu -o build/attn_tc_$name.o
objs=$(ls build/*.o | grep -v -E "build/attn_tc(_[a-z0-9]+)?\.o$")
nvcc -shared -o libo2b200_$name.so $objs build/attn_tc_$name.o -gencode arch=compute_100a,code=sm_100a -lcudart_static -Xcompiler -fPIC
echo built libo2b200_$name.so
