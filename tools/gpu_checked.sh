#!/bin/bash
# The memory-safety pass that can be run on this pool (compute-sanitizer is closed there): the library built with
# -DRM_CHECKED (device asserts on every queue index, tile index, pixel address, stack depth and exchange parameter) and
# the whole GPU suite run with it; then the production build again.   tools/gpu_checked.sh TAG
TAG=${1:-x}
RM_NVCC_EXTRA="-DRM_CHECKED" python -m rusty_marcher_b200.build --force > /dev/null 2>&1 || { echo "checked build failed"; exit 1; }
echo "checked build: $(/usr/local/cuda/bin/cuobjdump -elf rusty_marcher_b200/librm_b200.so 2>/dev/null | grep -c 'externs:.*__assertfail') kernels / functions reference __assertfail" | tee gpurun_out/checked_$TAG.txt
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q -rs >> gpurun_out/checked_$TAG.txt 2>&1; echo "pytest (RM_CHECKED) rc=$?" | tee -a gpurun_out/checked_$TAG.txt
tail -4 gpurun_out/checked_$TAG.txt
python -m rusty_marcher_b200.build --force > /dev/null 2>&1; echo "production build restored rc=$?"
