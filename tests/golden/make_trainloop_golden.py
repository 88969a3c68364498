#!/usr/bin/env python
"""Golden traces of the reference's epoch loop, early stopping and test procedures (train.py:147-255),
produced by THAT code: the statements from `min_loss = None` to the end of `/root/reference/train.py` are cut
out with `ast` and executed unmodified against recording stand-ins for the Keras model, the reader, h5py and
`keras.models.load_model` (none of which is installable here).

    python tests/golden/make_trainloop_golden.py       # needs /root/reference; writes tests/golden/trainloop.json

A trace is the sequence of calls the script makes: generators requested (set, sparsity, flags), fit / evaluate
step counts, models saved and reloaded, batches predicted, and the manual RMSE it prints for scripted
predictions. `tests/test_train_host.py` replays the same scripts through `train.run`."""
import ast
import contextlib
import io
import json
import os
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def loop_code():
    with open("/root/reference/train.py") as f:
        tree = ast.parse(f.read())
    start = [k for k, n in enumerate(tree.body) if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "min_loss"][0]
    return compile(ast.Module(body=tree.body[start:], type_ignores=[]), "train.py", "exec")


class Gen(object):
    """A Python-2 style generator object (`.next()`, train.py:233) over scripted batches."""

    def __init__(self, batches):
        self.it = iter(batches)

    def next(self):
        return next(self.it)

    __next__ = next


def run_case(script, patience, eval_mode, sizes, B, test_sparsities, saved_models_load=True):
    trace = []
    rs = np.random.RandomState(len(script) * 7 + patience)

    class Reader(object):
        train_set_size, val_set_size, test_set_size = sizes

        def data_gen(self, batch_size, data_sparsity, train_val_test="train", shuffle=True, auxilliary_mask_type="dropout",
                     aux_var_value=-1, return_target_count=False, sparse_representation=False,
                     pass_through_input_training=False):
            trace.append(["data_gen", train_val_test, data_sparsity, bool(return_target_count), bool(pass_through_input_training)])
            n = sizes[{"train": 0, "valid": 1, "test": 2}[train_val_test]] // batch_size
            batches = []
            for _ in range(n):
                t = np.round(rs.random_sample((batch_size, 5)) * 5)
                batches.append(([t * 0], t, int(np.count_nonzero(t))) if return_target_count else ([t * 0], t))
            return Gen(batches)

    class Model(object):
        metrics_names = ["loss", "mean_absolute_error", "accurate_MAE", "nMAE", "accurate_RMSE", "accurate_MSE"]

        def __init__(self, tag):
            self.tag, self.epoch = tag, 0

        def fit_generator(self, gen, steps, validation_data=None, validation_steps=None):
            trace.append(["fit_generator", float(steps), float(validation_steps)])
            v = script[self.epoch]
            self.epoch += 1
            return types.SimpleNamespace(history={"val_accurate_MSE": [v], "loss": [1.0]})

        def save(self, path):
            trace.append(["save", self.tag, path[len("models/NAME"):]])

        def evaluate_generator(self, gen, steps):
            trace.append(["evaluate_generator", self.tag, float(steps)])
            return [0.1, 0.2, 0.3, 0.4, 0.5, 0.6]

        def predict(self, input_list, batch_size=None, verbose=0):
            trace.append(["predict", self.tag])
            return input_list[0] + 1.0                    # every prediction is 1: SSE = sum((1 - t)^2)

    saved = set()
    m = Model("live")
    real_save = m.save

    def save_and_remember(path):
        saved.add(path)
        real_save(path)

    m.save = save_and_remember

    def load_model(path, custom_objects=None):
        if not saved_models_load or path not in saved:
            raise IOError("no such file " + path)
        trace.append(["load_model", path[len("models/NAME"):]])
        return Model("loaded" + path[len("models/NAME"):])

    class H5File(object):
        def __init__(self, *a):
            raise IOError("h5py is not available")

    ns = dict(np=np, m=m, data_reader=Reader(), keras=types.SimpleNamespace(models=types.SimpleNamespace(load_model=load_model)),
              h5py=types.SimpleNamespace(File=H5File), max_epochs=len(script), batch_size=B, train_sparsity=[1.0, 1.0],
              shuffle_data_every_epoch=True, auxilliary_mask_type=None, aux_var_value=-1, use_sparse_representation=False,
              pass_through_input_training=True, early_stopping_metric="val_accurate_MSE", patience=patience,
              model_save_path="models/", model_save_name="NAME", eval_mode=eval_mode, test_sparsities=test_sparsities,
              accurate_MAE=None, accurate_RMSE=None, nMAE=None, accurate_MSE=None)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        exec(loop_code(), ns)
    res = {"trace": trace, "best_epoch": ns["best_epoch"], "epochs_run": ns["i"] + 1 if eval_mode != "ablation" else None,
           "val_history": ns["val_history"], "tested": ns["best_m"].tag}
    if eval_mode == "fixed_split":
        res["manual_rmse"] = float(ns["RMSE"])
        res["ratings_count"] = int(ns["ratings_count"])
    # epochs_run in ablation mode: the test loop reuses `i`; count the fit calls instead
    res["epochs_run"] = sum(1 for e in trace if e[0] == "fit_generator")
    return res


def main():
    cases = []
    for script, patience in (([5, 4, 4.5, 4.2, 3.9, 4.0, 4.1, 4.3, 9, 9], 1), ([5, 6, 7, 8], 0), ([5, 6, 4, 6, 6, 6, 3, 9, 9, 9, 9], 2),
                             ([3, 2, 1, 0.5], 0), ([2, 2, 2, 2, 2], 1), ([7], 0), ([5, 4, 4, 3, 3], 0)):
        for eval_mode in ("fixed_split", "ablation"):
            sizes, B = (61, 23, 19), 4
            args = dict(script=script, patience=patience, eval_mode=eval_mode, sizes=sizes, B=B,
                        test_sparsities=[0.0, 0.5, 0.9])
            cases.append(dict(args, result=run_case(**args)))
    with open(os.path.join(HERE, "trainloop.json"), "w") as f:
        json.dump(cases, f, indent=0)
    for c in cases[:4]:
        print(c["script"], c["patience"], c["eval_mode"], {k: v for k, v in c["result"].items() if k != "trace"})
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
