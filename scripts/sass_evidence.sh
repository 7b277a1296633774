#!/bin/bash
# Static evidence from the built library (no GPU needed): registers / spills of every kernel and the
# SASS mnemonic histogram of the three dominant kernels.  What to look for is B200_PROFILING.md's list:
# UBLKCP = cp.async.bulk (TMA engine, 1-D), SYNCS = mbarrier, ACQBULK/ elect = bulk-copy issue by one
# elected thread, FFMA2/FADD2/FMUL2 = packed f32x2 arithmetic, no scalar FFMA in the sample path
# (every product and sum of the reference is rounded separately).
# Usage: bash scripts/sass_evidence.sh > profiles/r02_sass_evidence.txt
LIB=webradio_b200/libwebradio_b200.so
echo "cuobjdump -res-usage $LIB  (sm_100a; STACK/LOCAL 0 = no spills -- every kernel but demod_audio_kernel_v2<192>, whose STACK:8 is one"
echo "8-byte slot: a single STL/LDL pair around the out-of-line slow path of the FM discriminator's IEEE division)"
echo
cuobjdump -res-usage $LIB 2>/dev/null | grep -A1 "Function" | grep -v "^--" | paste - - \
  | sed -E 's/^ *Function ([^:]+):\s*/\1 /; s/ (TEXTURE|SURFACE|SAMPLER):0//g' | while read -r name rest; do
    printf "%-86s %s\n" "$(echo "$name" | c++filt | sed -E 's/\(anonymous namespace\):://g; s/\(.*//; s/void //' | cut -c1-86)" "$rest"
  done | sort
hist() {
  echo
  echo "SASS mnemonics of $2"
  cuobjdump -sass -fun "$1" $LIB 2>/dev/null | grep -E '^\s+/\*[0-9a-f]{4}\*/' \
    | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+ +//' | awk '{print $1}' | sed 's/[.;].*//' \
    | sort | uniq -c | sort -rn | awk '{printf "%6d %-14s", $1, $2; if (NR % 5 == 0) printf "\n"} END {printf "\n"}'
}
hist _ZN3wrd14chan_kernel_v3ILi127ELi50ELi2ELb0EEEvNS_8ChanArgsENS_6V3ArgsE "chan_kernel_v3<127,50,2,false> (cfg2: 64 receivers on one tuner)"
hist _ZN3wrd14chan_kernel_v4ILi255ELi50ELi8ELi3EEEvNS_8ChanArgsENS_6V4ArgsE "chan_kernel_v4<255,50,8,3> (cfg3: 1024 independent streams; LDGSTS = cp.async)"
hist _ZN3wrd14chan_kernel_v3ILi255ELi50ELi2ELb0EEEvNS_8ChanArgsENS_6V3ArgsE "chan_kernel_v3<255,50,2,false> (cfg3 fed raw bytes)"
hist _ZN3wrd21demod_audio_kernel_v2ILi192EEEvNS_14DemodAudioArgsE "demod_audio_kernel_v2<192>"
SPEC=$(cuobjdump -res-usage $LIB 2>/dev/null | grep -o '_ZN[A-Za-z0-9_]*spectrum_kernel_v2ILi32E[A-Za-z0-9_]*' | head -1)
hist "$SPEC" "spectrum_kernel_v2<32> (8192-point transforms; rows that straddle the carry buffer)"
SPEC3=$(cuobjdump -res-usage $LIB 2>/dev/null | grep -o '_ZN[A-Za-z0-9_]*spectrum_kernel_v3ILi32ELi2E[A-Za-z0-9_]*' | head -1)
hist "$SPEC3" "spectrum_kernel_v3<32,2> (cfg4: 8192-point transforms, hop 4096; UBLKCP = cp.async.bulk, SYNCS = mbarrier)"
