#!/bin/bash
DPM_LIB=$PWD/deeppointmap_b200/libdpm_prof.so python tools/gemm_profile.py 2>&1 | tail -30
