#!/bin/bash
# build kernel variants (experiments) into /root/repo/variants/ and time each on the GPU box
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
build() { # name, extra flags
  SMCRT_NVCC_EXTRA="$2" python -c "from rsmcrt_b200 import build as b; from pathlib import Path; b.build(force=True, out=Path('variants/$1.so'))"
}
"$@"
