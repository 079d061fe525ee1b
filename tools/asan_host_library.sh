#!/bin/bash
# The WHOLE library's host code (solver.cu's host side included: C ABI entry points, frame program, distributed layout and its
# symbolic replay) under AddressSanitizer + UBSan: nvcc builds a second .so with the sanitizers on the host compiler, the
# host-only tests of the CPU suite run against it (preloaded runtimes, ctypes).  ~2.5 min of nvcc + ~1.5 min of tests; no GPU.
# The default CPU suite covers plan.cpp / ingest.cpp the same way through a plain g++ harness (tests/test_sanitizers.py).
set -e
if [ -n "$SKIP_BUILD" ] && [ -f "${1:-/tmp/libsb_asan.so}" ]; then skip=1; fi
root=$(cd "$(dirname "$0")/.." && pwd); out=${1:-/tmp/libsb_asan.so}
cd "$root/softbodyunity_b200/csrc"
[ -n "$skip" ] || /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -ccbin /usr/bin/g++ \
  -Xcompiler -fPIC,-ffp-contract=off,-march=x86-64-v3,-pthread,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer \
  -shared -cudart static -o "$out" solver.cu plan.cpp ingest.cpp -lpthread
cd "$root"
drv=$(mktemp /tmp/asan_driver_XXXX.py)   # (a file, not stdin: the gloo tests spawn processes that re-import the main module)
cat > "$drv" <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import softbodyunity_b200.build as b
b.LIB = os.environ["SB_ASAN_LIB"]; b.stale = lambda: False
if __name__ == "__main__":
    import pytest
    sys.exit(pytest.main(["tests/test_plan_host.py", "tests/test_ingest.py", "tests/test_golden.py", "tests/test_shard_gloo.py",
                          "-q", "-m", "not gpu", "-p", "no:cacheprovider", "-s"]))
PY
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0:abort_on_error=0 \
UBSAN_OPTIONS=print_stacktrace=1 SB_ASAN_LIB="$out" python "$drv" > /tmp/asan_host_library.log 2>&1 || true
rm -f "$drv"
tail -2 /tmp/asan_host_library.log
echo "sanitizer reports: $(grep -c 'runtime error\|AddressSanitizer' /tmp/asan_host_library.log)  (log: /tmp/asan_host_library.log)"

# the CPU oracle (test infrastructure) under the same sanitizers: its known-answer tests, the golden fixtures and the planner tests that drive it
orc=/tmp/liborc_asan.so
gcc -O1 -g -std=c11 -fPIC -shared -fopenmp -ffp-contract=off -fno-fast-math -march=x86-64-v3 -fsanitize=address,undefined -fno-omit-frame-pointer \
  -o "$orc" oracle/xpbd_oracle.c -lm
drv=$(mktemp /tmp/asan_driver_XXXX.py)
cat > "$drv" <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import oracle.xpbd_oracle as o
o.LIB = "/tmp/liborc_asan.so"; o.build = lambda force=False: o.LIB
if __name__ == "__main__":
    import pytest
    sys.exit(pytest.main(["tests/test_oracle_kat.py", "tests/test_golden.py", "tests/test_plan_host.py", "-q", "-m", "not gpu", "-p", "no:cacheprovider", "-s"]))
PY
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0:abort_on_error=0 \
UBSAN_OPTIONS=print_stacktrace=1 python "$drv" > /tmp/asan_oracle.log 2>&1 || true
rm -f "$drv" "$orc"
tail -1 /tmp/asan_oracle.log
echo "oracle: sanitizer reports: $(grep -c 'runtime error\|AddressSanitizer' /tmp/asan_oracle.log)  (log: /tmp/asan_oracle.log)"
