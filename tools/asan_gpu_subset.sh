#!/bin/bash
# Host-side AddressSanitizer / UBSan pass over a subset of the GPU tests.
#   here:   python p4-phylogenetics_b200/_build.py -Xcompiler -fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer,-g
#           in a scratch copy of the repository; put its libp4b200.so and _pfhot*.so under build/asan/ (git-ignored, travels with gpurun)
#   box:    gpurun -- 'ASAN_TESTS="tests/test_gpu_newt.py ..." bash tools/asan_gpu_subset.sh'
# The CPU suite under the same build: LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" python -m pytest tests -m "not gpu"
cp build/asan/libp4b200.so build/asan/_pfhot*.so p4-phylogenetics_b200/
ASAN=$(gcc -print-file-name=libasan.so); UBSAN=$(gcc -print-file-name=libubsan.so)
export LD_PRELOAD="$ASAN $UBSAN"
export ASAN_OPTIONS=detect_leaks=0:halt_on_error=0:protect_shadow_gap=0:log_path=gpurun_out/asan_gpu
export UBSAN_OPTIONS=print_stacktrace=1:log_path=gpurun_out/ubsan_gpu
timeout ${ASAN_T:-110} python -m pytest ${ASAN_TESTS:-tests/test_gpu_parity.py tests/test_gpu_mcmc.py tests/test_gpu_fused20.py} -m gpu -q -p no:cacheprovider --durations=5 > gpurun_out/asan_gpu_tests${ASAN_TAG}.log 2>&1
tail -12 gpurun_out/asan_gpu_tests${ASAN_TAG}.log
ls gpurun_out | grep -c "asan_gpu\.\|ubsan_gpu\."
