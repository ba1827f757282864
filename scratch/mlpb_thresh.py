import os, sys
sys.path.insert(0, 'scratch')
from mlpb_time import run
T = int(os.environ.get('TT', '32'))
for (S, W, d) in ((32, 32, 2), (20, 25, 1)):
    for B in (1024, 2048, 4096, 8192, 16384, 65536):
        run(S, W, d, B, T, {'CC_B200_MLPB': '0'}, reps=2)
        run(S, W, d, B, T, {'CC_B200_MLPB': '1'}, reps=2)
