#!/bin/bash
# epilogue ablation of the narrow dilated layers (cta_group::1): what bounds the epilogue
S="8,64,64,4096,2;8,96,128,2048,4;8,128,256,512,16"
export AID_TC2_CG2=0
echo "== full";            TC_SHAPES=$S timeout 120 python tools/time_conv.py 3 2>&1 | tail -3
echo "== no residual";     NOR=1 TC_SHAPES=$S timeout 120 python tools/time_conv.py 3 2>&1 | tail -3
echo "== no stats";        NOSTATS=1 TC_SHAPES=$S timeout 120 python tools/time_conv.py 3 2>&1 | tail -3
echo "== no stores (256)"; AID_TC_DEBUG=256 TC_SHAPES=$S timeout 120 python tools/time_conv.py 3 2>&1 | tail -3
echo "== no tmem ld (512)"; AID_TC_DEBUG=512 TC_SHAPES=$S timeout 120 python tools/time_conv.py 3 2>&1 | tail -3
echo "== no stores, no residual"; NOR=1 AID_TC_DEBUG=256 TC_SHAPES=$S timeout 120 python tools/time_conv.py 3 2>&1 | tail -3
echo "== no stores, no residual, no stats"; NOSTATS=1 NOR=1 AID_TC_DEBUG=256 TC_SHAPES=$S timeout 120 python tools/time_conv.py 3 2>&1 | tail -3
echo "== epilogue body off (1)"; AID_TC_DEBUG=1 TC_SHAPES=$S timeout 120 python tools/time_conv.py 3 2>&1 | tail -3
echo "== no A/B loads (12): epilogue + MMA only"; AID_TC_DEBUG=12 TC_SHAPES=$S timeout 120 python tools/time_conv.py 3 2>&1 | tail -3
