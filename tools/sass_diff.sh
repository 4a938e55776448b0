#!/bin/bash
# Which kernels of the library differ, at the SASS level, between a reference commit and the working tree?
#   tools/sass_diff.sh <commit>
# Used when opt-in flavours are added without a GPU at hand: the kernels on the default path must stay bit-identical to
# the build that last passed the full GPU suite.  Kernels are matched by name without the anonymous-namespace hash;
# renamed template instantiations show up as NEW and have to be compared by hand (cuobjdump -sass -fun ...).
set -e
ref=${1:?usage: tools/sass_diff.sh <commit>}
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
git -C "$root" archive "$ref" utmos_b200/csrc include | tar -x -C "$tmp"
flags="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
mkdir -p "$tmp/old" "$tmp/new"
for f in ingest select tail mgpu convert synth; do
  (cd "$tmp/utmos_b200/csrc" && /usr/local/cuda/bin/nvcc $flags -I ../../include -I . -c $f.cu -o "$tmp/old/$f.o") &
  (cd "$root/utmos_b200/csrc" && /usr/local/cuda/bin/nvcc $flags -I ../../include -I . -c $f.cu -o "$tmp/new/$f.o") &
done
wait
clean() { grep -v "Function :\|identifier =" | sed 's#/\*[0-9a-f]\{4\}\*/##' | grep -v "^\s*/\* 0x"; }
short() { sed 's/_GLOBAL__N__[0-9a-f]*_[0-9]*_\([a-z]*\)_cu_[0-9a-f]\{8\}/\1/'; }
for f in ingest select tail mgpu convert synth; do
  cuobjdump -sass "$tmp/old/$f.o" | grep "Function :" | sed 's/.*Function : //' > "$tmp/o.txt"
  cuobjdump -sass "$tmp/new/$f.o" | grep "Function :" | sed 's/.*Function : //' > "$tmp/n.txt"
  for fn in $(cat "$tmp/n.txt"); do
    s=$(echo "$fn" | short); old=""
    for c in $(cat "$tmp/o.txt"); do [ "$(echo "$c" | short)" = "$s" ] && old=$c; done
    if [ -z "$old" ]; then echo "$f NEW  $(echo "$s" | c++filt | cut -c1-110)"; continue; fi
    a=$(cuobjdump -sass -fun "$old" "$tmp/old/$f.o" | clean | md5sum)
    b=$(cuobjdump -sass -fun "$fn" "$tmp/new/$f.o" | clean | md5sum)
    [ "$a" = "$b" ] || echo "$f DIFF $(echo "$s" | c++filt | cut -c1-110)"
  done
done
rm -rf "$tmp"
