#!/bin/bash
# SASS listings of the hot kernels of the built library for profiles/ (cuobjdump -sass, instruction encodings stripped) and the
# per-mnemonic static counts.  Usage: tools/sass_listing.sh <tag>      (reproduces from f16_mpc_oop_py_b200/libf16_b200.so)
TAG=${1:-r02}
SO=f16_mpc_oop_py_b200/libf16_b200.so
list() {  # <out-name> <mangled-name substring>
  cuobjdump -sass "$SO" 2>/dev/null | awk -v k="$2" '
    /Function :/ { on = index($0, k) > 0; if (on) print }
    on && /^ +\/\*[0-9a-f]+\*\// { sub(/ *\/\* 0x[0-9a-f]+ \*\/ *$/, ""); print }' > profiles/${TAG}_sass_$1.txt
  { echo "# static SASS instruction mix of $2 (loop bodies count once)"; bash tools/sass_mix.sh "$SO" "$2"; } > profiles/${TAG}_sass_$1_mix.txt
  wc -l profiles/${TAG}_sass_$1.txt
}
list step_hifi_fast_chunked step_hifi_fast_chunked_kernelILb0ELi0E
list step_hifi_fast_plain step_hifi_fast_kernelILb1ELb0ELi384ELi0E
list step_hifi_fast_lqr_mpc_columns step_hifi_fast_kernelILb1ELb1ELi384ELi14880664E
list linearise_fast_hifi linearise_fast_kernelILi1ELi256ELb1E
list linearise_fast_hifi_forward linearise_fast_kernelILi1ELi384ELb0E
list xdot_fast_hifi_calc_xdot xdot_fast_kernelILi1ELb0ELi256E
list xdot_fast_hifi_nlplant xdot_fast_kernelILi1ELb1ELi384E
list stats_partial stats14partial_kernel
