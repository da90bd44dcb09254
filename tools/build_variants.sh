#!/bin/bash
# build tuning variants of libbrs.so: tools/build_variants.sh name1="flags" name2="flags" ...  (fails if any build fails)
mkdir -p build/var; rm -f build/var/*.so
pids=()
for spec in "$@"; do
  name=${spec%%=*}; flags=${spec#*=}
  ( nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 $flags -o build/var/libbrs_$name.so aitk.robots_b200/csrc/brs_api.cu 2>&1 | grep -E "error" ; test -f build/var/libbrs_$name.so ) &
  pids+=($!)
done
rc=0; for p in "${pids[@]}"; do wait $p || rc=1; done
ls build/var; exit $rc
