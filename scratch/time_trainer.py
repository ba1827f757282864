"""scratch: wall time per ModelControllerTrainer.step() of the public mirrors at the reference's own sizes (configs[0], configs[1])."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import OrderedDict
import numpy as np, torch
from chain_control_b200.examples.neural_ode_controller_compact_example import make_neural_ode_controller
from chain_control_b200.examples.neural_ode_model_compact_example import make_neural_ode_model
from chain_control_b200.train import (ModelControllerTrainer, SupervisedDataset, TrainingOptionsController, TrainingOptionsModel,
                                      UnsupervisedDataset, make_dataloader, optim)
rng = np.random.default_rng(0)
def bench(tr, n=30):
    for _ in range(5): tr.step()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(n): tr.step()
    torch.cuda.synchronize(); return 1e3 * (time.time() - t0) / n
model = make_neural_ode_model(1, OrderedDict(y=1), 0.01, 50, f_depth=0, u_transform=np.arctan)
with torch.no_grad():
    W = model.f[0].weight; W[:, :50] -= torch.eye(50, device=W.device)
u = rng.uniform(-1, 1, size=(8, 1000, 1)).astype(np.float32)
with torch.no_grad(): y = model.unroll_batched(u)
mo = TrainingOptionsModel(make_dataloader(SupervisedDataset(u, y), key=2, n_minibatches=1), optim.chain(optim.clip_by_global_norm(1.0), optim.adam(3e-3)))
print("configs[0] ModelTrainer 8 x 1000: %.2f ms per trainer.step()" % bench(ModelControllerTrainer(model, model_train_options=mo)))
refs = OrderedDict(y=(np.sign(rng.normal(size=(64, 1, 1))) * rng.uniform(0, 3, size=(64, 1, 1)) * np.ones((64, 1001, 1))).astype(np.float32))
ctrl = make_neural_ode_controller(2, 1, 0.01, 5, f_depth=0)
co = TrainingOptionsController(make_dataloader(UnsupervisedDataset(refs), key=1, n_minibatches=1), optim.chain(optim.clip_by_global_norm(1.0), optim.adam(1e-3)))
print("configs[1] ControllerTrainer 64 x 1001: %.2f ms per trainer.step()" % bench(ModelControllerTrainer(model, ctrl, controller_train_options=co)))
