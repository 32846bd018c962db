#!/bin/bash
# Checked build: libthetarrt with device-side bounds assertions (-DTRRT_CHECKED, trrt_device.cuh) into
# profiles/tools/_variants/checked.so; run the GPU suite against it with
#     python -m pytest tests -m gpu -q --trrt-so profiles/tools/_variants/checked.so
set -e
cd "$(dirname "$0")/../.."
python profiles/tools/variant_bench.py build checked -DTRRT_CHECKED
