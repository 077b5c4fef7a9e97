#!/bin/sh
# Stage the reference's own CPU implementation under oracle/_ref/.
#
# The reference is Fortran 2003 and this image has no Fortran compiler (gfortran,
# flang, nvfortran, ifort: all absent), so it cannot be rebuilt from source.  The
# repository ships the author's own `gfortran -fopenmp` build (`prealpha`, GCC 14.2,
# -O0); it runs here once the single missing versioned symbol
# (_gfortran_os_error_at@GFORTRAN_10) is supplied by a shim libgfortran.so.5 that
# forwards everything else to the libgfortran vendored inside scipy.
#
# Outputs only into oracle/_ref/ (binary + shim).  Nothing is copied into history.
set -e
here=$(cd "$(dirname "$0")" && pwd)
ref=${PREALPHA_REFERENCE:-/root/reference}
out="$here/_ref"
if [ ! -f "$ref/prealpha" ]; then
    if [ -x "$out/prealpha" ]; then echo "oracle/_ref: using prebuilt files (no reference checkout here)"; exit 0; fi
    echo "oracle/_ref: reference checkout not found at $ref and nothing prebuilt" >&2; exit 0
fi
SC=$(python3 -c 'import scipy, os; print(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs"))')
GF=$(ls "$SC"/libgfortran-*.so.5.0.0 | head -1)
for cand in "$SC"/libgfortran-*.so.5.0.0; do
    # prefer the copy that exports the GFORTRAN_8 version node used by the binary
    if objdump -T "$cand" 2>/dev/null | grep -q GFORTRAN_8; then GF=$cand; break; fi
done
mkdir -p "$out/lib"
cat > "$out/shim.map" <<'MAP'
GFORTRAN_8 { };
GFORTRAN_10 { global: _gfortran_os_error_at; } GFORTRAN_8;
MAP
cat > "$out/shim.c" <<'SHIM'
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
void _gfortran_os_error_at(const char *where, const char *msg, ...) {
  va_list ap; va_start(ap, msg); fprintf(stderr, "%s: ", where); vfprintf(stderr, msg, ap);
  fputc('\n', stderr); va_end(ap); abort(); }
SHIM
gcc -shared -fPIC "$out/shim.c" -Wl,--version-script="$out/shim.map" -Wl,-soname,libgfortran.so.5 \
    -o "$out/lib/libgfortran.so.5" -Wl,--no-as-needed "$GF" -Wl,-rpath,"$SC"
cp "$ref/prealpha" "$out/prealpha"
chmod +x "$out/prealpha"
echo "$SC" > "$out/scipy_libs_path"
echo "oracle/_ref: staged $out/prealpha (+ libgfortran shim -> $GF)"
