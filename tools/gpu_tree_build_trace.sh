#!/bin/bash
# where the time of a 2^20-primitive upload goes (WRT_TRACE_BUILD=1), host build against device build
set -u
mkdir -p gpurun_out
WRT_TRACE_BUILD=1 python - > gpurun_out/r02_build_trace.txt 2>&1 <<'P'
import sys, importlib, time, os
sys.path.insert(0,'.'); sys.path.insert(0,'oracle')
wrt = importlib.import_module("zig-weekend-raytracer_b200")
import wro_py as wro
sc = wro.OracleScene("synthetic", seed=1, n_prims=1<<20)
flat = sc.flatten()
for dev in (-1, 0, 0):
    i,_,_ = wrt.build_trees(flat, dev, records=False)
    print("wrt_build_trees device", dev, "build", round(i.build_ms,2), "total", round(i.total_ms,1), flush=True)
for mode in ("0", "1", "1"):
    os.environ["WRT_DEVICE_BUILD"] = mode
    with wrt.Context(0) as c:
        t=time.time(); c.upload_scene(flat); dt=(time.time()-t)*1e3
        st=c.stats()
        print("upload_scene WRT_DEVICE_BUILD", mode, "wall", round(dt,1), "upload_ms", round(st.upload_ms,1), "tree_build_ms", round(st.tree_build_ms,2), "on device", st.tree_build_device, flush=True)
P
echo "rc=$?"; cat gpurun_out/r02_build_trace.txt
