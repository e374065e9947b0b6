#!/bin/bash
# SASS evidence for the Blackwell-native claim: mnemonic counts of the built library, total and per kernel.
# usage: bash tools/sass_table.sh > profiles/rNN_sass_table.txt
lib=nerf-sos_b200/lib/libnerfsos.so
tmp=$(mktemp)
cuobjdump -sass $lib > $tmp
echo "# SASS evidence: cuobjdump -sass $lib | grep -c <mnemonic>   (nvcc $(nvcc --version | grep -o 'release [0-9.]*' | cut -d' ' -f2), -gencode arch=compute_100a,code=sm_100a)"
for m in UTCHMMA LDTM STTM UBLKCP LDGSTS UTCBAR UTMALDG UTMASTG HGMMA "SYNCS.PHASECHK" "UTCATOMSWS\|UTCALLOC"; do printf "%-16s %s\n" "$m" "$(grep -c "$m" $tmp)"; done
echo
echo "# per kernel (UTCHMMA / LDTM / STTM / UBLKCP):"
awk '/Function : /{name=$3} /UTCHMMA/{a[name]++} /LDTM/{b[name]++} /STTM/{c[name]++} /UBLKCP/{d[name]++} END{for(n in a) printf "%5d %5d %5d %5d  %s\n", a[n], b[n]+0, c[n]+0, d[n]+0, n}' $tmp | sed 's/_ZN4nsos[0-9]*_GLOBAL__N__[0-9a-f]*_[0-9]*_[a-z_]*_cu_[0-9a-f]*[0-9]*//' | sort -k5
rm -f $tmp
