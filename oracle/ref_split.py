"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's offline splitter,
`/root/reference/TrainValidTestSplit.py`, without pandas.

Only `tests/` may import this module; nothing under `omnidirectional_collaborative_filtering_b200/` does.

Parity status: PINNED. `tests/golden/make_split_golden.py` runs the reference's own script (unmodified but
for its parameter lines) in the build container on small CSVs and stores every file it writes under
`tests/golden/split/`; `tests/test_split_files.py` checks this restatement against those files byte for
byte. Where the script cannot finish under Python 3 (json.dump of np.int64: all-integer rows, integer
timestamps next to string ids, `save_users_and_items`) this writes what Python 2 wrote - integers as
integers - which is the declared behaviour of the product's splitter too; those cases are pinned only
through the CSV files the script did write before it died.

What is restated (reference file:line):
  read_ratings      - `pd.read_csv` + the per-column dtype inference it applies (:34)
  row_value         - what `ratings.iloc[i][col]` yields after the row is upcast to one dtype (:124-125)
  build_user_item_dict - :121-149     map_inputs_to_targets - :183-195     merge_timestamps - :197-211
  split_data        - :31-118, build_and_save :159-181, convert_and_save_mml :213-219
"""
from __future__ import annotations

import csv
import json

import numpy as np


def _infer(fields):
    """(kind, values): 'int' / 'float' / 'str' column the way read_csv types it."""
    try:
        ints = [int(f) for f in fields]
        if all(-2 ** 63 <= v < 2 ** 63 for v in ints):      # beyond int64 pandas leaves int64 too (uint64 / object):
            return "int", ints                               # declared here as a float64 column
    except ValueError:
        pass
    try:
        return "float", [float(f) if f != "" else float("nan") for f in fields]
    except ValueError:
        return "str", list(fields)


def read_ratings(path, n_columns):
    with open(path, newline="", encoding="utf-8") as f:
        rows = [r for r in csv.reader(f) if r]
    body = rows[1:]
    for r in body:
        if len(r) != n_columns:
            raise ValueError("expected %d fields, got %d" % (n_columns, len(r)))
    cols = [_infer([r[c] for r in body]) for c in range(n_columns)]
    kinds = [k for k, _ in cols]
    row_kind = "str" if "str" in kinds else ("float" if "float" in kinds else "int")
    return cols, row_kind, len(body)


def row_value(cols, row_kind, c, i):
    kind, vals = cols[c]
    if kind == "str":
        return vals[i]
    if row_kind == "float" or kind == "float":
        return float(vals[i])
    return int(vals[i])


def _csv_field(cols, c, i):
    kind, vals = cols[c]
    v = vals[i]
    if kind == "int":
        return str(v)
    if kind == "float":
        return "" if v != v else repr(v)
    if any(ch in v for ch in ',"\r\n'):
        return '"' + v.replace('"', '""') + '"'
    return v


def split_data(full_data_filepath, output_filepath, schema_type="movielens", trainvalidtest_split=(.8, .1, .1),
               build_data_for_omni=True, include_timestamps=True, save_users_and_items=False,
               reverse_user_item_data=False, rng=np.random):
    if reverse_user_item_data:                                            # :27-29
        output_filepath = output_filepath + "reverse_item-user/"
    cols, row_kind, n = read_ratings(full_data_filepath, 3 if schema_type == "netflix" else 4)
    user_c, item_c = (1, 0) if reverse_user_item_data else (0, 1)         # :40-69
    cast_user_to_int = schema_type == "movielens"
    order = rng.permutation(n)                                            # :74
    n_tr = int(n * trainvalidtest_split[0])
    n_va = int(n * trainvalidtest_split[1])
    tr, va, te = order[:n_tr], order[n_tr:n_tr + n_va], order[n_tr + n_va:]
    te_in = order[:n_tr + n_va]                                           # :83
    suffix = "_withtimestamps" if include_timestamps else ""

    def save_mml(idx, name):                                              # :213-219
        with open(output_filepath + name, "w", encoding="utf-8", newline="") as f:
            for i in idx:
                fields = [_csv_field(cols, user_c, i), _csv_field(cols, item_c, i), _csv_field(cols, 2, i)]
                if include_timestamps:
                    fields.append(_csv_field(cols, 3, i))
                f.write(",".join(fields) + "\n")

    save_mml(te_in, "train_data_mml" + suffix + ".csv")                   # :91-96
    save_mml(te, "test_data_mml" + suffix + ".csv")

    def build_user_item_dict(idx):                                        # :121-149
        ratings, stamps = {}, {}
        for i in idx:
            u = row_value(cols, row_kind, user_c, i)
            user = str(int(u)) if cast_user_to_int else (repr(u) if isinstance(u, float) else str(u))
            item = row_value(cols, row_kind, item_c, i)
            rating = row_value(cols, row_kind, 2, i)
            stamp = row_value(cols, row_kind, 3, i) if include_timestamps else None
            ratings.setdefault(user, []).append((item, rating))
            stamps.setdefault(user, []).append((item, stamp))
        return ratings, stamps

    def map_inputs_to_targets(input_set, targets):                        # :183-195
        return {user: (input_set[user] if user in input_set else None) for user in targets}

    def merge_timestamps(ins, outs):                                      # :197-211 (on copies: no aliasing visible)
        merged = {user: list(l) for user, l in ins.items()}
        for user, l in outs.items():
            if user in merged:
                merged[user].extend(l)
            else:
                merged[user] = list(l)
        return merged

    def save(obj, which):                                                 # :151-157
        with open(output_filepath + "ratingsByUser_dicts" + suffix + "_" + which + ".json", "w") as f:
            json.dump(obj, f)

    if build_data_for_omni:                                               # :98-103, :159-181
        train = build_user_item_dict(tr)
        save((train[0], train[1]) if include_timestamps else train[0], "train")
        for which, tg_idx, in_dicts in (("valid", va, train), ("test", te, None)):
            if in_dicts is None:
                in_dicts = build_user_item_dict(te_in)
            out = build_user_item_dict(tg_idx)
            paired = (map_inputs_to_targets(in_dicts[0], out[0]), out[0])
            save((paired, merge_timestamps(in_dicts[1], out[1])) if include_timestamps else paired, which)

    if save_users_and_items:                                              # :105-118
        def unique(c):
            seen, out = set(), []
            for v in cols[c][1]:
                k = "nan" if v != v else v
                if k not in seen:
                    seen.add(k)
                    out.append(v)
            return out
        with open(output_filepath + "unique_items_list.json", "w") as f:
            json.dump(unique(item_c), f)
        users = unique(user_c)
        as_str = [str(int(x)) for x in users] if cast_user_to_int else [repr(x) if isinstance(x, float) else str(x) for x in users]
        with open(output_filepath + "unique_users_list.json", "w") as f:
            json.dump(as_str, f)
    return output_filepath
