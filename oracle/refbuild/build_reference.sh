#!/bin/bash
# ORACLE INFRASTRUCTURE (never used by the product path).
#
# Builds an importable copy of the UNMODIFIED reference sources (Python+Cython)
# in a scratch directory, against the serial stand-ins under stubs/ for the
# dependencies that are absent from this image (mpi4py/MPI, modepy, h5py,
# METIS).  /root/reference is read-only, so the sources are copied first.
# The result is used by oracle/refbuild/make_golden.py to generate the
# fixtures in tests/golden/ and (optionally) as the CPU baseline.
#
# usage: build_reference.sh [SRC=/root/reference] [DST=/tmp/refbuild]
set -e
SRC=${1:-/root/reference}
DST=${2:-/tmp/refbuild}
HERE=$(cd "$(dirname "$0")" && pwd)
JOBS=${JOBS:-8}

mkdir -p "$DST"
rm -rf "$DST/src" "$DST/stubs"
mkdir -p "$DST/src"
cp -r "$SRC"/{packageTools,base,fem,multilevelSolver,nl,PyNucleus,VERSION,config.yaml} "$DST/src/"
cp -r "$HERE/stubs" "$DST/stubs"
# the modepy stand-in loads oracle/triangle_rules.py relative to the repo root
REPO_ROOT=$(cd "$HERE/../.." && pwd)
sed -i "s#^_root = .*#_root = '$REPO_ROOT'#" "$DST/stubs/modepy/__init__.py"

cat > "$DST/src/config.yaml" <<EOC
compiler_c: gcc
compiler_c++: g++
arch: detect
mpi: generic
compileArgs: [-O3, -pipe, -Wno-cpp, -w]
includeDirs: []
linkArgs: [-O3, -pipe]
cythonDirectives:
  binding: true
  embedsignature: true
  language_level: '2'
setupProfiling: false
annotate: false
threads: $JOBS
use_ccache: false
use_cholmod: false
EOC

export PYTHONPATH="$DST/stubs:$DST/src/packageTools:$DST/src/base:$DST/src/fem:$DST/src/multilevelSolver:$DST/src/nl:$DST/src:$PYTHONPATH"
export PYNUCLEUS_BUILD_PARALLELISM=$JOBS

(cd "$DST/stubs/mpi4py/.." && python mpi4py/setup.py build_ext --inplace > "$DST/build_mpi4py.log" 2>&1)
(cd "$DST/stubs" && python PyNucleus_metisCy/setup.py build_ext --inplace > "$DST/build_metis.log" 2>&1)

for pkg in base fem multilevelSolver nl; do
    echo "=== building $pkg ($(date +%T))"
    (cd "$DST/src/$pkg" && python setup.py build_ext --inplace > "$DST/build_$pkg.log" 2>&1) || { echo "FAILED: $pkg (see $DST/build_$pkg.log)"; tail -30 "$DST/build_$pkg.log"; exit 1; }
done
echo "=== done ($(date +%T)); use:"
echo "export PYTHONPATH=$DST/stubs:$DST/src/packageTools:$DST/src/base:$DST/src/fem:$DST/src/multilevelSolver:$DST/src/nl:$DST/src"

# ---- install the runtime files (stripped .so + .py only) into oracle/_ref ----
# oracle/_ref is git-ignored (build output, not source) but travels to the GPU
# box, where it serves as the CPU baseline (bench.py --impl reference).
REF="$REPO_ROOT/oracle/_ref"
rm -rf "$REF"; mkdir -p "$REF"
for pkg in base/PyNucleus_base fem/PyNucleus_fem multilevelSolver/PyNucleus_multilevelSolver nl/PyNucleus_nl packageTools/PyNucleus_packageTools; do
    name=$(basename $pkg)
    mkdir -p "$REF/$name"
    (cd "$DST/src/$pkg" && find . \( -name '*.so' -o -name '*.py' \) -exec cp --parents {} "$REF/$name" \;)
done
for st in mpi4py h5py modepy PyNucleus_metisCy; do
    mkdir -p "$REF/$st"
    (cd "$DST/stubs/$st" && find . \( -name '*.so' -o -name '*.py' \) -exec cp --parents {} "$REF/$st" \;)
done
# inside oracle/_ref the modepy stand-in finds oracle/triangle_rules.py two levels up
sed -i "s#^_root = .*#_root = os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..', '..'))#" "$REF/modepy/__init__.py"
find "$REF" -name '*.so' -exec strip --strip-debug {} \;
du -sh "$REF"
