#!/usr/bin/env python
"""The module-level parameters of the reference's scripts, read out of their source with `ast` (never executed):
`train.py:22-59` (+ the save-name expression and its two suffixes, :59,76-80) and `TrainValidTestSplit.py:17-25`.

    python tests/golden/make_params_golden.py      # needs /root/reference; writes tests/golden/script_params.json

`tests/test_train_host.py` holds `train.TrainConfig` and `splitter.split_data` to these names and defaults."""
import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def top_level_assignments(path, stop_at=None):
    with open(path) as f:
        tree = ast.parse(f.read())
    out = {}
    for node in tree.body:
        if stop_at and getattr(node, "lineno", 0) > stop_at:
            break
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
            try:
                out[name] = {"value": ast.literal_eval(node.value)}
            except ValueError:
                out[name] = {"source": ast.unparse(node.value)}
    return out


def signatures(path, wanted):
    """{qualified name: [[parameter, default or "<required>"], ...]} of the listed functions / methods."""
    with open(path) as f:
        tree = ast.parse(f.read())
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef):
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and (node.name + "." + fn.name) in wanted:
                    a = fn.args
                    names = [x.arg for x in a.args]
                    defaults = ["<required>"] * (len(names) - len(a.defaults)) + [ast.literal_eval(d) for d in a.defaults]
                    out[node.name + "." + fn.name] = [[n, d] for n, d in zip(names, defaults) if n != "self"]
    return out


def main():
    train = top_level_assignments("/root/reference/train.py", stop_at=60)
    # the expression building model_save_name, evaluated with the defaults (str() of lists etc. included)
    env = {k: v["value"] for k, v in train.items() if "value" in v}
    name = eval(train["model_save_name"]["source"], {}, dict(env))
    if env["reverse_user_item_data"]:
        name += "_itemUserReverse"                      # train.py:76-78
    name += "_" + env["dataset"] + "_"                  # train.py:80 (the timestamp follows)
    split = top_level_assignments("/root/reference/TrainValidTestSplit.py", stop_at=26)
    sigs = signatures("/root/reference/data_reader.py", {"data_reader.__init__", "data_reader.data_gen", "data_reader.split_for_validation"})
    sigs.update(signatures("/root/reference/model.py", {"omni_model.__init__", "omni_model.save_weights", "omni_model.load_weights",
                                                        "omni_model.replace_dense_layer_weights", "omni_model.manually_load_all_weights",
                                                        "omni_model.make_trainable", "omni_model.load_and_fix_for_denoising_autoencoders"}))
    with open(os.path.join(HERE, "script_params.json"), "w") as f:
        json.dump({"train": train, "default_model_save_name_prefix": name, "split": split, "signatures": sigs}, f, indent=1)
    print(sorted(train), name, sorted(split), sep="\n")


if __name__ == "__main__":
    main()
