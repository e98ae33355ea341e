#!/bin/bash
# Sanitizer passes over the kernel bodies.
#   scripts/sanitize.sh ubsan   CPU: the emulation build (same kernel bodies + the host side of the C ABI) compiled with
#                               -fsanitize=undefined, whole emulation test file under halt_on_error (no GPU needed)
#   scripts/sanitize.sh order   CPU: the emulation suite with the threads of a CTA run in reverse and in pseudo-random order
#                               between barriers (LITHO_EMU_ORDER): a missing barrier makes the result order-dependent
#   scripts/sanitize.sh gpu     B200: compute-sanitizer memcheck / racecheck / synccheck on the small GPU parity cases
#                               (shared-memory exchange, TMA tile buffer + mbarrier, named barriers, the tcgen05
#                               direct-solver kernel, the two-stage rim sums)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$ROOT/lithographysimulator_b200/csrc
case "${1:-ubsan}" in
ubsan)
    OUT=${TMPDIR:-/tmp}/litho_ubsan; mkdir -p $OUT
    FLAGS="-O1 -g -std=c++17 -fPIC -DLITHO_EMU -I$ROOT/tests/emu -I$CSRC -fsanitize=undefined -fno-sanitize-recover=undefined -x c++"
    (for m in 16 32 64 128 256 512 1024 2048 4096 8192 16384; do
        echo "g++ $FLAGS -DLITHO_INST_M=$m -c $CSRC/kernels_inst.cu -o $OUT/inst_$m.o"; done
     echo "g++ $FLAGS -c $CSRC/litho_abi.cu -o $OUT/abi.o") | xargs -P 8 -I{} sh -c "{}"
    g++ -shared -fPIC -fsanitize=undefined -o $OUT/liblitho_emu.so $OUT/*.o
    make -C $CSRC emu -j8 > /dev/null
    cp $ROOT/tests/emu/liblitho_emu.so $OUT/plain.so
    trap 'cp $OUT/plain.so $ROOT/tests/emu/liblitho_emu.so' EXIT
    cp $OUT/liblitho_emu.so $ROOT/tests/emu/liblitho_emu.so
    LD_PRELOAD=$(g++ -print-file-name=libubsan.so) UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
        python -m pytest $ROOT/tests/test_emu_kernels.py -x -q
    ;;
order)
    for o in reverse random; do
        LITHO_EMU_ORDER=$o python -m pytest $ROOT/tests/test_emu_kernels.py -x -q
    done
    ;;
gpu)
    for tool in memcheck racecheck synccheck; do
        compute-sanitizer --tool $tool --error-exitcode 9 \
            python -m pytest $ROOT/tests/test_gpu_parity.py -m gpu -x -q \
            -k "(golden and (demo64 or np2_96 or wrap_128 or direct_64)) or (tensor_core_path and (64 or 128)) or input_validation" \
            || { echo "compute-sanitizer $tool reported errors"; exit 9; }
    done
    ;;
esac
