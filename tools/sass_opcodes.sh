#!/bin/bash
# Opcode evidence for the shipped library: TMA (UTMALDG/UTMAPF/UBLKCP), mbarrier (SYNCS.*), setmaxnreg (USETMAXREG),
# packed FP32x2 (FFMA2/FADD2/FMUL2), L2 discard (CCTL.*) and the memory opcodes, per kernel family.
#   bash tools/sass_opcodes.sh > profiles/r02_sass_opcodes.txt
cd "$(dirname "$0")/.."
LIB=fourier_b200/lib/libfourier.so.0.1.0
echo "cuobjdump -sass $LIB ($(stat -c %s $LIB) bytes, $(cuobjdump -lelf $LIB | wc -l) cubin(s) for $(cuobjdump -lelf $LIB | grep -o 'sm_[0-9a]*' | sort -u | tr '\n' ' '))"
echo "whole library:"
cuobjdump -sass $LIB | grep -oE '\b(UTMALDG|UTMAPF|UBLKCP|UTMASTG|SYNCS|USETMAXREG|FFMA2|FADD2|FMUL2|DFMA|DADD|DMUL|CCTL|LDGSTS|LDS|STS|LDG|STG|ATOMS|ATOMG|RED|BAR|MUFU|SHFL|HMMA|UTCMMA|WGMMA)[.A-Za-z0-9_]*' | sed -E 's/(\.E|\.STRONG|\.GPU|\.SYS|\.CONSTANT)//g' | sort | uniq -c | sort -rn | head -60
for k in fused_twopass_kernel bluestein_fused_kernel onchip_fft_kernel tile_kernel cta_fft_kernel rows_exchange_kernel column_kernel 15exchange_kernel; do
  echo; echo "kernels matching $k: $(cuobjdump -sass $LIB | grep -c "Function : .*$k")"
  cuobjdump -sass $LIB | awk -v k="$k" '/Function : /{p = ($0 ~ k)} p' | grep -oE '\b(UTMALDG|UTMAPF|UBLKCP|SYNCS|USETMAXREG|FFMA2|FADD2|FMUL2|DFMA|CCTL|LDS|STS|LDG|STG)[.A-Za-z0-9_]*' | sed -E 's/(\.E|\.STRONG|\.GPU|\.SYS|\.CONSTANT)//g' | sort | uniq -c | sort -rn | head -16
done
