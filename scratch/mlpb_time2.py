import os, sys
sys.path.insert(0, 'scratch')
from mlpb_time import run
T = int(os.environ.get('TT', '32'))
for nh in ('2', '4'):
    run(64, 64, 2, 18944, T, {'CC_B200_MLPB_NH': nh})
    run(64, 64, 2, 65536, T, {'CC_B200_MLPB_NH': nh})
    run(64, 64, 1, 65536, T, {'CC_B200_MLPB_NH': nh})
    run(32, 32, 2, 65536, T, {'CC_B200_MLPB_NH': nh, 'CC_B200_MLPB': '1'})
