#!/bin/bash
# Builds a kernel-experiment variant of the library: tools/build_variant.sh NAME [-DFLAG=VALUE ...]
# -> ros2-recursive-patchwork-implementation_b200/_variants/NAME.so  (timed by tools/gpu_variants.py)
exec python "$(dirname "$0")/../ros2-recursive-patchwork-implementation_b200/_build.py" "$@" 2>&1 | grep -v "warning #177\|declared but never\|constexpr int\|\^\|^$\|Remark"
