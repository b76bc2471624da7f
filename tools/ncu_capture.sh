#!/bin/sh
# `ncu --set full` captures of the three tensor-core kernels on representative layers (run under gpurun, one GPU),
# then summarised here with profiles/summarize_ncu.py.  The .ncu-rep files stay in gpurun_out/ (scratch).
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
# forward conv 256->256 3x3 on 64x128x64 (2-CTA cluster kernel), full epilogue
MICRO_SHAPES=1 MICRO_MODES=full timeout 200 $N -k regex:conv_umma -s 3 -c 1 -f -o gpurun_out/conv_fwd_256 python tests/bench_conv_micro.py pair > gpurun_out/ncu_a.log 2>&1
# forward conv 128->128 3x3 on 64x128x64 (single-CTA kernel, wide-B form), full epilogue
MICRO_SHAPES=0 MICRO_MODES=full timeout 200 $N -k regex:conv_umma -s 3 -c 1 -f -o gpurun_out/conv_fwd_128 python tests/bench_conv_micro.py single > gpurun_out/ncu_b.log 2>&1
# filter gradient 256->256 3x3 on 64x128x64 (2-CTA cluster kernel)
MICRO_SHAPE=4 MICRO_VARIANT=wgrad MICRO_ITERS=1 timeout 200 $N -k regex:wgrad_umma -c 1 -f -o gpurun_out/wgrad_256 python tests/bench_dgrad_micro.py > gpurun_out/ncu_c.log 2>&1
# stride-2 data gradient, four parity classes in one launch, out + masked + colsum epilogue
MICRO_SHAPE=0 MICRO_VARIANT=out+masked+colsum MICRO_ITERS=1 timeout 200 $N -k regex:conv_umma -c 1 -f -o gpurun_out/dgrad_s2 python tests/bench_dgrad_micro.py > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep
