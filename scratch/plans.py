import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
os.environ['CC_B200_VERBOSE'] = '1'
import ctypes as C
from ccspecs import model_variants, controller_variants
from cchelpers import conv
import emu_lib as E
L = E.emu()
m = conv(model_variants()['compact_s50_d0']).to_c(); c = conv(controller_variants()['compact_s5_d0']).to_c()
for B in [64, 4096, 65536]:
    print('B', B, 'open ws', L.cc_rollout_workspace_bytes(C.byref(m), B, 1000, 1), 'closed ws', L.cc_closedloop_workspace_bytes(C.byref(c), C.byref(m), B, 1001, 1))
    print(' supported fwd', L.cc_module_supported(C.byref(m), 0))
