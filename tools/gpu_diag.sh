#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nproc | tee gpurun_out/nproc.txt
echo "=== tma"; timeout 30 ./tools/diag_tma 2>&1 | tail -5 | tee gpurun_out/diag_tma.log
echo "=== step"; timeout 150 python -u -X faulthandler - > gpurun_out/diag_step.log 2>&1 <<'PY'
import faulthandler, sys, time
faulthandler.dump_traceback_later(100, exit=True)
t0=time.time()
def say(*a): print("[%.1f]"%(time.time()-t0), *a, flush=True)
say("start")
sys.path.insert(0,'tests'); sys.path.insert(0,'oracle')
import numpy as np
import __graft_entry__ as g
say("entry imported")
P = g.load_package(); say("pkg loaded")
L = P.lib(); say("lib loaded")
ctx = P.Context(0); say("ctx created")
import parity_util as pu; say("parity_util imported")
case = pu.Case(dims=(6,5,4)); say("case built")
mesh = case.box.make_mesh(ctx, tile_nodes=32); say("mesh created", mesh.stats())
pu.upload_state(P, mesh, case); ctx.sync(); say("state uploaded")
x = mesh.download("velocity"); say("download ok", float(abs(x - case.fields["velocity"].reshape(x.shape)).max()))
mesh.mdot_edge(1.0, 1.0); say("mdot launched"); ctx.sync(); say("mdot synced")
mesh.peclet_edge("viscosity", P.peclet_fn("classic",1.0)); ctx.sync(); say("peclet synced")
mesh.register("dpdx_out", P.NW_NODE, 3); mesh.nodal_grad_edge("pressure","dpdx_out"); ctx.sync(); say("grad synced")
ls = P.LinearSystem(mesh, P.NW_LINSYS_HYPRE, 1); ls.buildEdgeToNodeGraph(); ls.finalizeLinearSystem(); say("linsys finalized")
ls.zeroSystem(); ls.assemble_continuity_edge(**pu.CONT_OPTS); say("continuity launched"); ctx.sync(); say("continuity synced")
say(pu.run_lowmach_case(P, ctx, dims=(6,5,4), tile_nodes=32))
say(pu.run_lowmach_case(P, ctx, dims=(12,10,8), tile_nodes=64))
say("done")
PY
tail -30 gpurun_out/diag_step.log
