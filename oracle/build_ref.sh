#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.
# Lifts the reference's CPU verifier loops out of /root/reference/main.mm where
# they lie (nothing is copied into the tracked tree: the cut lines land in
# oracle/_ref/, which is git-ignored) and compiles them, wrapped by
# oracle/ref_shim.cpp, into oracle/_ref/libref_cpu.so.
# The reference's own build (clang++ -fobjc-arc + Metal frameworks, Makefile:1-16)
# cannot run on Linux; only these plain-C++ line ranges of main() are built.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${REFERENCE_DIR:-/root/reference}"
src="$ref/main.mm"
out="$here/_ref"
if [ ! -f "$src" ]; then
  echo "build_ref: $src not present; keeping any prebuilt $out/libref_cpu.so" >&2
  exit 0
fi
mkdir -p "$out"
cut_lines() { sed -n "$1,$2p" "$src" > "$out/$3"; }
cut_lines 24 30 ref_init_random.inc      # initRandom
cut_lines 128 159 ref_forward.inc        # non-causal CPU reference
cut_lines 550 578 ref_causal.inc         # causal CPU reference
cut_lines 1092 1179 ref_backward.inc     # backward CPU reference
# Guard against a reference whose line numbers moved: each cut must start and
# end where the survey says it does.
head -1 "$out/ref_init_random.inc" | grep -q 'void initRandom'
head -1 "$out/ref_forward.inc" | grep -q 'for (int i = 0; i < N; ++i)'
head -1 "$out/ref_causal.inc" | grep -q 'std::vector<float> O_ref'
head -1 "$out/ref_backward.inc" | grep -q 'std::vector<float> P_cpu'
tail -8 "$out/ref_backward.inc" | grep -q "dK_cpu\[j \* D + d\] +="
g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -I"$here" \
    -o "$out/libref_cpu.so" "$here/ref_shim.cpp"
rm -f "$out"/*.inc   # the cut lines are build intermediates; only the .so is kept
echo "build_ref: built $out/libref_cpu.so from $src"
