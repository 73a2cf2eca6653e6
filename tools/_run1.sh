PBX_SO=pixelbox_b200/lib/exp/lib_bprof.so python tools/batch_prof.py 10000000 256 8 | tail -4
PBX_SO=pixelbox_b200/lib/exp/lib_bprof.so python tools/batch_prof.py 10000000 256 1024 | tail -4
