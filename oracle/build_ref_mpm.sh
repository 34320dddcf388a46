#!/bin/sh
# Builds oracle/_ref/libref_mpm_{snow,fc,jelly}.so (MaterialModel = MMSnow / MMFixedCorotated / MMJelly) from the reference sources where they lie under $1.
# Only line ranges are extracted, into a temporary directory that is removed afterwards: nothing from
# the reference is copied into the repo or left in _ref/ (which holds the .so files only).
set -e
REF="$1"; CXX="$2"; FLAGS="$3"
cd "$(dirname "$0")"
mkdir -p _ref
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
sed -n '18,53p' "$REF/src/linalg.cu" > "$TMP/gen_linalg.inc"          # device half of linalg
sed -n '6,8p;14,178p' "$REF/src/mpm.cu" > "$TMP/gen_kernels.inc"      # consts + the three kernels
INC="-I$TMP -Ishim -I$REF/include -I/usr/local/cuda/include"
W="-Wno-unused-variable -Wno-sign-compare -Wno-endif-labels -Wno-attributes -Wno-unused-but-set-variable -Wno-unused-function -w"
$CXX $FLAGS $W $INC -fvisibility=hidden -o _ref/libref_mpm_snow.so ref_mpm_host.cpp
$CXX $FLAGS $W $INC -fvisibility=hidden -DREF_FIXED_COROTATED -o _ref/libref_mpm_fc.so ref_mpm_host.cpp
$CXX $FLAGS $W $INC -fvisibility=hidden -DREF_JELLY -o _ref/libref_mpm_jelly.so ref_mpm_host.cpp
