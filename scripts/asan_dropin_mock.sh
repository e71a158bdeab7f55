#!/bin/bash
# AddressSanitizer run of the C drop-in dispatcher on CPU: the reference engine (system allocator instead of its own heap, so
# ASan sees every block) + voxplat_b200/host/vp_chunkset_manage.c + the host mock of vp_multi, driven by the tests of
# tests/test_dropin_mock.py (the concurrent edit / dispatch / acknowledge test N times).  Needs /root/reference.
# usage: bash scripts/asan_dropin_mock.sh [repetitions]
set -e
cd "$(dirname "$0")/../oracle"
REF=${REF:-/root/reference}; OUT=${OUT:-/tmp/vp_asan}; mkdir -p $OUT
F="-std=gnu99 -O1 -g -fno-omit-frame-pointer -fsanitize=address -DUSE_SYSTEM_ALLOC -fcommon -fopenmp -w -fPIC -I$REF/src -I$REF/include"
gcc $F -Igfx_shim -c $REF/src/gfx/vsplat.c -o $OUT/vsplat.o
gcc $F -Igfx_shim -c gfx_shim/gl_capture.c -o $OUT/gl_capture.o
gcc $F -Iref_shim -Dchunkset_manage=chunkset_manage_cpu -c $REF/src/chunkset.c -o $OUT/chunkset_cpu.o
gcc $F -Iref_shim -DVR_WITH_GPU -I../include -shared -o $OUT/libvoxref_mock_asan.so $OUT/chunkset_cpu.o $REF/src/chunkset/mesher.c $REF/src/chunkset/rle.c \
    $REF/src/chunkset/edit.c $REF/src/mem.c $REF/src/event.c ref_shim/shim.c ref_harness.c $OUT/vsplat.o $OUT/gl_capture.o \
    ../voxplat_b200/host/vp_chunkset_manage.c mock/vp_multi_mock.c vox_oracle.c -lm -lpthread
cat > $OUT/run.py <<PY
import sys, ctypes as C
sys.path.insert(0, "$PWD/.."); sys.path.insert(0, "$PWD/../tests")
import test_dropin_mock as T
lib = C.CDLL("$OUT/libvoxref_mock_asan.so"); lib.vr_world_create.restype = C.c_void_p; lib.vr_init(C.c_uint64(2 << 30)); lib.vr_set_scratch_scale(12)
for name in ("test_dropin_bookkeeping_matches_the_reference_dispatcher", "test_rle_only_chunks_are_uploaded_as_streams", "test_edit_between_the_two_passes_is_not_lost"):
    getattr(T, name)(lib); print(name, "ok", flush=True)
for rep in range(${1:-5}):
    T.test_no_edit_is_lost_while_the_dispatcher_runs(lib); print("concurrent run", rep, "ok", flush=True)
PY
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python $OUT/run.py 2>&1 | grep -v "^chunkset.c\|^mem.c"
