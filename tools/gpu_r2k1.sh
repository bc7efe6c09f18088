#!/bin/bash
# r2k1 (1 GPU): `ncu --set full` of K1 / dgrad pack / K1b / gate backward on the 512 -> 512 layer (the HBM-bound case).
O=gpurun_out; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1 REPMODE_NO_BUILD=1
timeout 110 ncu --set full --clock-control none -k regex:'reparam_fwd_rows_kernel|pack_dgrad_h16_kernel|reparam_bwd_reg_kernel|gate_bwd_block_kernel' \
  -c 12 -o $O/r2k1_full_k1 -f python tools/ncu_k1_512.py > $O/r2k1_ncu.log 2>&1
echo "ncu exit $?"; tail -2 $O/r2k1_ncu.log | cut -c1-200
[ -s $O/r2k1_full_k1.ncu-rep ] && timeout 40 ncu -i $O/r2k1_full_k1.ncu-rep --page raw --csv > $O/r2k1_full_k1_raw.csv 2>/dev/null
ls -la $O/r2k1_full_k1* | cut -c20-
rm -f $O/r2k1_full_k1.ncu-rep
echo done
