#!/bin/bash
# compute-sanitizer memcheck over the kernels added at the end of round 2: out block, in-conversion upsampling, thin-in CO4, init block
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_out_block.py -m gpu -q -k "case0 or case2 or case4 or case6" 2>&1 | grep "ERROR SUMMARY\|passed\|failed\|Invalid" | head
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_unet.py -m gpu -q -k "paper_network_config1 or per_sample_sigma" 2>&1 | grep "ERROR SUMMARY\|passed\|failed\|Invalid" | head
