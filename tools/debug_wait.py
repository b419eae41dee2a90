import sys, ctypes
sys.path.insert(0, '.')
import numpy as np, torch
from tsp_gnn_b200 import instances as inst, params as P, _lib
from tsp_gnn_b200.engine import Engine
EV, W, C, y, nv, ne = inst.synth_batch([40] * 128, seed=42)
eng = Engine(64, "bf16x3", 0)
eng.set_params(P.init_params(64, seed=0))
eng.set_option("fused", 1)
eng.plan(nv, ne, EV.src, EV.dst)
dW = torch.from_numpy(W.astype(np.float32).reshape(-1)).cuda(); dC = torch.from_numpy(C.astype(np.float32).reshape(-1)).cuda()
eng.init_embeddings(dW, dC)
eng.step(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 16)()
fn = _lib.lib.tspgnn_debug_wait_info
fn.argtypes = [ctypes.c_void_p]; fn.restype = ctypes.c_int
print("rc", fn(buf))
print("timeouts:", buf[0])
BAR = 0x38000
names = {0:"w",1:"x_full",2:"x_empty",3:"h_full0",4:"h_full1",5:"h_full2",6:"acc_full0",7:"acc_full1",8:"act_ready0",9:"act_ready1",10:"spare",11:"p_w",12:"p_x_full",13:"p_h_full0",14:"p_h_full1",15:"p_h_full2",16:"boot0",17:"boot1"}
for i in range(1, min(16, int(buf[0]) + 1)):
    v = buf[i]; addr = (v >> 32) & 0xFFFFF; par = (v >> 31) & 1; blk = (v >> 12) & 0xFFF; tid = v & 0xFFF
    print("  addr 0x%x parity %d block %d warp %d lane %d" % (addr, par, blk, tid >> 5, tid & 31))
