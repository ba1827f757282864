import os, sys
sys.path.insert(0, 'scratch')
from mlpb_time import run
T = int(os.environ.get('TT', '32'))
for (S, W, d) in ((5, 10, 2), (16, 16, 2), (12, 16, 1)):
    for B in (1024, 4096, 16384, 65536, 262144):
        run(S, W, d, B, T, {'CC_B200_MLPB': '0'}, reps=3)
        run(S, W, d, B, T, {'CC_B200_MLPB': '1'}, reps=3)
