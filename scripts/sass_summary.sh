#!/bin/bash
# SASS evidence for the design claims: which instruction families the sm_100a cubins of the product library contain.
#   bash scripts/sass_summary.sh > profiles/sass_summary.md        (needs only cuobjdump: runs in the GPU-less container)
LIB=${1:-hexed_b200/libhexed_b200.so}
T=$(mktemp)
cuobjdump -sass "$LIB" > "$T" 2>/dev/null
echo "# SASS summary of \`$LIB\`"
echo
echo "\`cuobjdump -sass $LIB | grep -c <mnemonic>\` over all $(grep -c 'Function :' "$T") kernels of the $(grep -c 'arch = sm_100a' "$T") sm_100a cubins ($(grep 'arch = ' "$T" | sort -u | tr '\n' ' '))."
echo
echo "| mnemonic | count | what it shows |"
echo "|---|---|---|"
row() { printf "| \`%s\` | %s | %s |\n" "$1" "$(grep -c -- "$1" "$T")" "$2"; }
row UBLKCP "1-D bulk TMA copies (cp.async.bulk): the staging of every pipelined kernel"
row SYNCS "mbarrier operations the bulk copies complete on"
row UTMALDG "tensor-map TMA loads: none -- every input of an element is ONE contiguous run, no tensor map needed"
row UTCMMA "tcgen05 MMA: none -- FP64 has no tcgen05 path"
row DMMA "FP64 mma.sync: none -- the kernels are HBM / shared-memory bound, not FP64-issue bound (north-star condition for DMMA not met)"
row DFMA "FP64 fused multiply-add: the arithmetic"
row DMUL "FP64 multiply"
row DADD "FP64 add"
row MUFU.RCP64H "FP64 reciprocal seeds (divisions by mass / determinant as reciprocal + multiply)"
row MUFU.RSQ64H "FP64 rsqrt seeds (sound speed in max_dt, characteristic BCs)"
row LDS.128 "128-bit shared loads (contiguous lines of the vector line map)"
row LDS.64 "64-bit shared loads"
row STS.128 "128-bit shared stores"
row STS.64 "64-bit shared stores"
row LDG.E.64 "64-bit global loads"
row LDG.E.128 "128-bit global loads"
row STG.E.64 "64-bit global stores"
row CCTL "L2 prefetches of the late inputs (prefetch.global.L2)"
row SHFL "warp shuffles (max_dt / residual reductions)"
row ATOMG "global atomics (running minimum of max_dt, admissibility flags)"
row BAR.SYNC "CTA barriers"
row LDL "local-memory loads (spills + dynamically indexed arrays of the generic any-row-size kernels)"
row STL "local-memory stores"
echo
echo "Per hot kernel (registers / spill from \`cuobjdump -res-usage\`):"
echo
echo '```'
cuobjdump -res-usage "$LIB" 2>/dev/null | grep -A1 -E "local_euler_pipe_kernel|local_euler_pipe2d|ns_local_line|neighbor_euler_kernelILi3ELi6|max_dt_euler_screen|ns_reconcile_bulk_kernel" | grep -E "Function|REG" | sed -e 's/Function \(.*\):/\1/' | paste - - | sed -e 's/\s\+/ /g' | cut -c1-220 | grep -E "ILi6E|ILi3ELi6E" | sort -u | head -40
echo '```'
rm -f "$T"
