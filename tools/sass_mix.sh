#!/bin/bash
# SASS instruction mix of one kernel of the built library: tools/sass_mix.sh <object-or-.so> <kernel-name-substring>
# (cuobjdump -sass; counts per mnemonic, registers from the ptxas logs under csrc/build/).  Static counts: loop bodies
# count once.
f=${1:-f16_mpc_oop_py_b200/libf16_b200.so}; k=${2:-step_hifi_fast_kernelILb1ELb0ELi384ELi0E}
cuobjdump -sass "$f" 2>/dev/null | awk -v k="$k" '
  /Function :/ { on = index($0, k) > 0 }
  on && /^ +\/\*[0-9a-f]+\*\// { m = $2; sub(/\..*/, "", m); sub(/;/, "", m); if (m ~ /^@/) { m = $3; sub(/\..*/, "", m); sub(/;/, "", m) } c[m]++; n++ }
  END { for (m in c) printf "%6d %s\n", c[m], m | "sort -rn"; close("sort -rn"); printf "%6d TOTAL\n", n }'
