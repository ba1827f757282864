"""timeline of one time step of the tcgen05 forward kernel (trace build of the library; scratch)"""
import sys, os, ctypes as C
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from chain_control_b200 import _lib
_lib.LIB_PATH = os.path.join('scratch', 'libccb200_trace.so')
_lib._lib = _lib.CcLib(_lib.LIB_PATH)
from oracle import np_oracle as O
from ccspecs import model_variants, controller_variants
from cchelpers import conv
from chain_control_b200 import ops
dev = 'cuda'
T_ = lambda a: torch.tensor(np.asarray(a, np.float32), device=dev)
ms_ = model_variants()['compact_s50_d0']; cs_ = controller_variants()['compact_s5_d0']
spec = conv(ms_); cs = conv(cs_)
theta = T_(O.init_params(ms_, 1, stable=True)); thc = T_(O.init_params(cs_, 2))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 37888
T = 200
u = torch.rand(B, T, 1, device=dev) * 2 - 1
for which in ('open', 'closed'):
    if which == 'open': ops.rollout_fwd(spec, theta, u)
    else: ops.closedloop_fwd(cs, spec, thc, theta, u)
    torch.cuda.synchronize()
    buf = (C.c_longlong * 512)()
    rc = _lib.lib()._dll.cc_tc_read_trace(buf)
    tr = np.array(buf[:]).reshape(8, 64)
    t0 = tr[0, 60]
    print('==', which, 'B', B, 'step start -> write_state+publish; per stage: [pub_start, arrived, a_ready(w0), issued(w0), d_ready wake]')
    for w in range(8):
        row = tr[w]
        s = 'warp%d: start %5d |' % (w, row[60] - t0)
        for st in range(4):
            r = row[8 * st: 8 * st + 5] - t0
            s += ' st%d pub %5d arr %5d' % (st, r[0], r[1]) + (' ardy %5d iss %5d' % (r[2], r[3]) if w % 4 == 0 else '') + ' wake %5d |' % r[4]
        s += ' final done %5d' % (row[61] - t0)
        print(s)
