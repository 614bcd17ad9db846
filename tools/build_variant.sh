#!/bin/bash
# Builds a second libb200w.so with extra -D flags on ONE source file (what-if / previous-revision variants for
# same-box A/B runs), from the objects of the current build:
#   tools/build_variant.sh <name> <source.cu> -DFLAG [-DFLAG2 ...]   ->  ab_prev/lib_<name>.so   (git-ignored, ships)
set -euo pipefail
NAME=$1; SRC=$2; shift 2
mkdir -p ab_prev/$NAME
python -c "import __graft_entry__ as g; g.build()" > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC \
  -Xcompiler -fvisibility=hidden "$@" -c runbooks_b200/csrc/$SRC -o ab_prev/$NAME/${SRC%.cu}.o
OBJS=""
for o in runbooks_b200/build/*.o; do
  if [ "$(basename $o)" = "${SRC%.cu}.o" ]; then OBJS="$OBJS ab_prev/$NAME/${SRC%.cu}.o"; else OBJS="$OBJS $o"; fi
done
nvcc -shared -o ab_prev/lib_$NAME.so $OBJS -ldl -Xcompiler -fPIC
ls -la ab_prev/lib_$NAME.so
