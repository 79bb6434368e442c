#!/usr/bin/env bash
# Builds oracle/_ref/libdegensac_ref.so from the UNMODIFIED reference sources where they lie
# (read-only /root/reference), following SURVEY.md §8(c).  Test infrastructure only.
# Output goes only into oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
SRC="$REF/src/pydegensac"
if [ ! -d "$SRC/degensac" ]; then
  echo "build_ref: $SRC not present (expected on the GPU box) - keeping any prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj_m" "$OUT/obj_d"
CFLAGS="-O3 -DNDEBUG -fPIC -fcommon -w -I$SRC -I$SRC/degensac"
# matutls archive list: matutls/CMakeLists.txt:12-26 (unitary.c/ortho.c reference a missing `unfl`:
# a static archive drops those unreferenced members, exactly like the reference's CMake build).
MATUTLS="atou1 cmprt hevmax mcopy qrbdv solvru trncm atovm csolv hmgen minv qrecvc solvtd trnm chouse cvmul house
 mmul qreval sv2u1v unitary chousv eigen housev ortho qrevec sv2uv utrncm cmattr eigval ldumat otrma qrlsq sv2val
 utrnhm cmcpy evmax ldvmat otrsm rmmult vmul cminv hconj lsqsv psinv ruinv svdu1v cmmul heigval matprt qrbdi smgen
 svduv cmmult heigvec mattr qrbdu1 solvps svdval matconsts"
for f in $MATUTLS; do
  [ -f "$SRC/matutls/$f.c" ] && gcc $CFLAGS -c "$SRC/matutls/$f.c" -o "$OUT/obj_m/$f.o" &
done
# degensac sources: CMakeLists.txt:28-41
for f in DegUtils exp_ranF exp_ranH Ftools hash Htools ranF ranH2el ranH rtools utools lapwrap; do
  gcc $CFLAGS -c "$SRC/degensac/$f.c" -o "$OUT/obj_d/$f.o" &
done
# second flavour of the two drivers with the reference's own compile-time option __FINAL_LSQ__ (exp_ranF.h:28-29,
# exp_ranH.c:16): the oracle of DGB200_FLAG_FINAL_LSQ
mkdir -p "$OUT/obj_l"
# (exp_ranH.c only: the F driver's __FINAL_LSQ__ text assigns a Score to an unsigned, exp_ranF.c:1702, and does not compile)
gcc $CFLAGS -D__FINAL_LSQ__ -c "$SRC/degensac/exp_ranH.c" -o "$OUT/obj_l/exp_ranH.o" &
wait
rm -f "$OUT/libmatutls.a" "$OUT/libdegensac_support.a" "$OUT/libdegensac_support_lsq.a"
ar rcs "$OUT/libmatutls.a" "$OUT"/obj_m/*.o
ar rcs "$OUT/libdegensac_support.a" "$OUT"/obj_d/*.o
cp "$OUT/obj_l/exp_ranH.o" "$OUT/obj_d/exp_ranH.o"
ar rcs "$OUT/libdegensac_support_lsq.a" "$OUT"/obj_d/*.o
# LAPACK (dsyev_/dgesvd_): third party, unpinned in the reference (CMakeLists.txt:6). Use the OpenBLAS 0.3.15
# that ships inside the opencv-python-headless wheel of this image.
SP="$(python -c 'import sysconfig; print(sysconfig.get_paths()["purelib"])')"
BLASDIR="$SP/opencv_python_headless.libs"
BLAS="$(ls "$BLASDIR"/libopenblas*.so 2>/dev/null | head -1 || true)"
if [ -z "$BLAS" ]; then echo "build_ref: no OpenBLAS found under $BLASDIR" >&2; exit 1; fi
gcc -O2 -fPIC -shared -o "$OUT/libdegensac_ref.so" "$HERE/ref_harness.c" \
  -Wl,--wrap=time,--wrap=srand,--wrap=rand,--wrap=random \
  -Wl,--whole-archive "$OUT/libdegensac_support.a" -Wl,--no-whole-archive "$OUT/libmatutls.a" \
  "$BLAS" -Wl,--disable-new-dtags,-rpath,"$BLASDIR" -lm
gcc -O2 -fPIC -shared -o "$OUT/libdegensac_ref_lsq.so" "$HERE/ref_harness.c" \
  -Wl,--wrap=time,--wrap=srand,--wrap=rand,--wrap=random \
  -Wl,--whole-archive "$OUT/libdegensac_support_lsq.a" -Wl,--no-whole-archive "$OUT/libmatutls.a" \
  "$BLAS" -Wl,--disable-new-dtags,-rpath,"$BLASDIR" -lm
rm -rf "$OUT/obj_m" "$OUT/obj_d" "$OUT/obj_l"
echo "build_ref: built $OUT/libdegensac_ref.so and libdegensac_ref_lsq.so (LAPACK: $BLAS)"
