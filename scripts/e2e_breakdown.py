"""Where does the host-visible time of one public-API train step go? (debug helper)"""
import ctypes as C
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from omnidirectional_collaborative_filtering_b200 import _lib, optimizers
from omnidirectional_collaborative_filtering_b200.data_reader import data_reader
from omnidirectional_collaborative_filtering_b200.model import omni_model

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "ml10m"]
fs = bench.make_dataset(w)
rd = data_reader(fs.n_cols, fs.train.n_rows, "", eval_mode="fixed_split", data=fs)
aux = w["aux"]
om = omni_model(w["layers"], w["hidden"], fs.n_cols, 128, dense_activation=w["act"], use_causal_info=aux is not None,
                use_both_masks=aux == "both", dropout_probability=w["dropout"], auxilliary_mask_type=aux)
m = om.model
m.compile(optimizers.Adagrad(lr=0.005), "mean_squared_error", rating_range=4.0)
g = rd.data_gen(128, w["sparsity"], "train", True, aux, -1, pass_through_input_training=w["pass_through"])
lib = _lib.lib()
for i in range(40):
    t0 = time.perf_counter(); b = next(g)
    t1 = time.perf_counter(); h = m._ensure(b.n_rows, b.n_entries, b.aux_type)
    t2 = time.perf_counter(); dev = b.upload(None)
    t3 = time.perf_counter()
    args = m._args(b); rec = np.empty(8, dtype=np.float32)
    _lib.check(lib.ocf_train_step(h, dev.handle, C.byref(args), _lib.ptr(rec), None))
    t4 = time.perf_counter()
    print("step %2d entries %7d gen %.2f ensure %.2f upload %.2f step+sync %.2f ms" %
          (i, b.n_entries, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3)))
