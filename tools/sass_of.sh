#!/bin/bash
# Developer tool: SASS of one kernel of vibrato_b200/csrc/kernels.o (substring of the mangled name) -> /tmp/k.sass
#   tools/sass_of.sh k_viterbi2ILi8ELi0ELb1
O=/root/repo/vibrato_b200/csrc/kernels.o
cuobjdump -sass $O | awk -v pat="$1" '/Function :/{f = index($0, pat) > 0} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" \
  | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*//' > /tmp/k.sass
wc -l /tmp/k.sass
