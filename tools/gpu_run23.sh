#!/bin/bash
# pure epilogue (no MMAs, no operand loads: AID_TC_DEBUG 2|4|8 = 14) of the N = 96 / N = 64 layers: what is it bound by
S="8,96,128,2048,4;8,64,64,4096,2"
for ew in 8 16; do
export AID_TC2_EW=$ew
echo "== EW $ew: epilogue only";           AID_TC_DEBUG=14 TC_SHAPES=$S timeout 120 python tools/time_conv.py 3 2>&1 | tail -2
echo "== EW $ew: no residual";  NOR=1 AID_TC_DEBUG=14 TC_SHAPES=$S timeout 120 python tools/time_conv.py 3 2>&1 | tail -2
echo "== EW $ew: no stats";  NOSTATS=1 AID_TC_DEBUG=14 TC_SHAPES=$S timeout 120 python tools/time_conv.py 3 2>&1 | tail -2
echo "== EW $ew: no stores";  AID_TC_DEBUG=270 TC_SHAPES=$S timeout 120 python tools/time_conv.py 3 2>&1 | tail -2
echo "== EW $ew: no tmem ld";  AID_TC_DEBUG=526 TC_SHAPES=$S timeout 120 python tools/time_conv.py 3 2>&1 | tail -2
echo "== EW $ew: no stores, no residual, no stats";  NOR=1 NOSTATS=1 AID_TC_DEBUG=270 TC_SHAPES=$S timeout 120 python tools/time_conv.py 3 2>&1 | tail -2
echo "== EW $ew: no residual, no stats";  NOR=1 NOSTATS=1 AID_TC_DEBUG=14 TC_SHAPES=$S timeout 120 python tools/time_conv.py 3 2>&1 | tail -2
done
