#!/bin/sh
# (dev) SASS evidence of the headline kernel for profiles/: opcode histogram, the asynchronous-copy / barrier / dependency
# instructions with their addresses, packed-arithmetic counts, resource usage.
#   sh tools/sass_listing.sh > profiles/rNN_sass_fused_static_fundus.txt
set -e
root="$(cd "$(dirname "$0")/.." && pwd)"
tmp="$(mktemp -d)"
cd "$tmp"
cuobjdump -xelf all "$root/ramdsir_b200/libramdsir.so" >/dev/null
k='_ZN7ramdsir23ram_fused_static_kernelINS_4SCfgILi256ELi256ELi25EEELb0ELb0EfNS_6EpiSatEEEvPKT2_PKfPKlPKiPfT3_NS_10MixScalarsES8_SD_'
cuobjdump -sass -fun "$k" ram_fused.sm_100a.cubin > k.sass
echo "# SASS of the headline kernel: ram_fused_static_kernel<SCfg<256,256,25>, BANK=0, DUAL=0, float, EpiSat> (sm_100a)"
echo "# cuobjdump -sass -fun <mangled> of ram_fused.sm_100a.cubin extracted from ramdsir_b200/libramdsir.so (tools/sass_listing.sh)"
echo
echo "## opcode histogram (static instruction counts)"
grep -E '^\s+/\*[0-9a-f]{4}\*/' k.sass | awk '{op=$2; if (op ~ /^@/) op=$3; sub(/\..*/, "", op); sub(/;/, "", op); print op}' | sort | uniq -c | sort -rn
echo
echo "## bulk-async copy (TMA engine, non-tensor form: cp.async.bulk -> UBLKCP), mbarrier (SYNCS), programmatic dependent launch (PREEXIT = launch_dependents, ACQBULK = griddepcontrol.wait), L2 prefetch (CCTL.E.PF2), tensor memory (UTCALLOC / STTM / LDTM = tcgen05.alloc / st / ld)"
grep -E 'UBLKCP|SYNCS|PREEXIT|ACQBULK|FENCE|CCTL|STTM|LDTM|UTC|TCGEN' k.sass | sed 's#/\* 0x[0-9a-f]* \*/##'
echo
printf "packed pairs: FADD2 %s  FMUL2 %s  FFMA2 %s   operand modifiers folded by ptxas: swapped halves (.LO_HI) %s, per-half negation (.NP/.PN) %s, scalar broadcast (.F32 source) %s\n" \
  "$(grep -c 'FADD2' k.sass)" "$(grep -c 'FMUL2' k.sass)" "$(grep -c 'FFMA2' k.sass)" "$(grep -c 'LO_HI' k.sass)" "$(grep -cE '\.(NP|PN)' k.sass)" "$(grep -cE 'R[0-9]+(\.reuse)?\.F32[ ,;]' k.sass)"
printf "FFMA.SAT %s   LDG.E.128 %s   STG.E.128 %s   LDS.128 %s   STS.128 %s   BAR.SYNC %s\n" \
  "$(grep -c 'FFMA.SAT' k.sass)" "$(grep -c 'LDG.E.128' k.sass)" "$(grep -c 'STG.E.128' k.sass)" "$(grep -c 'LDS.128' k.sass)" "$(grep -c 'STS.128' k.sass)" "$(grep -c 'BAR.SYNC' k.sass)"
echo
echo "## registers / shared memory (cuobjdump --dump-resource-usage)"
cuobjdump --dump-resource-usage ram_fused.sm_100a.cubin 2>/dev/null | paste - - | grep -E "static_kernel" | sed 's/.*Function \(.*\):.*\(REG:[0-9]*\) \(STACK:[0-9]*\).*/\2 \3 | \1/' | c++filt | sed 's/(.*//' | cut -c1-170
rm -rf "$tmp"
