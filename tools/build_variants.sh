#!/usr/bin/env bash
# A/B builds of libpolytope_b200.so with different -D flags, into polytope_b200/ab/ (git-ignored
# *.so, travels with gpurun).  Select one at run time with PB200_LIB=polytope_b200/ab/<name>.so.
#   tools/build_variants.sh name1 "-DFLAG=.." name2 "-DFLAG=.. -DFLAG2=.." ...
set -euo pipefail
cd "$(dirname "$0")/../polytope_b200/csrc"
mkdir -p ../ab
ARCH="-gencode arch=compute_100a,code=sm_100a"
# PB200_AB_ONLY="pb200" recompiles only those translation units per variant and links the others from
# the objects of the last `make` (the flags must then only matter to the recompiled units); variants
# are built in parallel.
ONLY=${PB200_AB_ONLY:-"pb200 sets hull diff peak lp_cta"}
build_one() {
    name=$1; flags=$2
    tmp=$(mktemp -d)
    for f in pb200 sets hull diff peak lp_cta; do
        if [[ " $ONLY " == *" $f "* ]]; then
            nvcc -O3 -lineinfo -std=c++17 $ARCH -Xcompiler -fPIC $flags -c -o $tmp/$f.o $f.cu &
        else
            cp $f.o $tmp/$f.o
        fi
    done
    wait
    nvcc $ARCH -shared -o ../ab/$name.so $tmp/pb200.o $tmp/sets.o $tmp/hull.o $tmp/diff.o $tmp/peak.o $tmp/lp_cta.o
    rm -rf $tmp
    echo "built ab/$name.so  ($flags)"
}
while [ $# -ge 2 ]; do
    build_one "$1" "$2" &
    shift 2
done
wait
