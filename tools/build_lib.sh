#!/bin/sh
# rebuild libdpig.so in-tree (same as __graft_entry__.build() without importing torch)
cd "$(dirname "$0")/.." && python -c "
import sys, importlib
sys.path.insert(0, '.')
print(importlib.import_module('disentangled-person-image-generation_b200.build').build())"
