"""The two file-format neighbours of the hot path, both native
(`csrc/ocf_etl.cpp`) and both checked against files the REFERENCE's own scripts wrote
(`tests/golden/make_split_golden.py`):

  * splitter: `splitter.split_data` writes the same bytes as `TrainValidTestSplit.py` for the same CSV and
    NumPy seed; `oracle/ref_split.py` (the pandas-free restatement) is pinned by the same files and then
    serves as the checker for the cases the reference cannot finish under Python 3 and for random CSVs;
  * ingest: `data_reader(..., use_json=True)` on the reference's files yields the batches the reference's
    `data_reader.py` yields from them, bit for bit; the native parser agrees with `json.load` + the
    per-rating Python path on every golden file and on adversarial JSON.
No GPU needed: these entry points are host code."""
import filecmp
import json
import os
import shutil

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import ref_split
from omnidirectional_collaborative_filtering_b200 import _lib, ingest, splitter
from omnidirectional_collaborative_filtering_b200.data_reader import _csr_from_lists, data_reader
from tests.conftest import GOLDEN
from tests.helpers import host_densify

SPLIT = os.path.join(GOLDEN, "split")
with open(os.path.join(SPLIT, "cases.json")) as _f:
    CASES = json.load(_f)


def _golden_dir(case):
    return os.path.join(SPLIT, case["name"], "reverse_item-user" if case["reverse_user_item_data"] else "")


def _run(fn, case, out, **kw):
    np.random.seed(case["seed"])
    return fn(os.path.join(SPLIT, case["name"], "ratings.csv"), str(out) + "/", case["schema_type"],
              include_timestamps=case["include_timestamps"], reverse_user_item_data=case["reverse_user_item_data"], **kw)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
@pytest.mark.parametrize("impl", ["native", "oracle"])
def test_splitter_writes_the_reference_bytes(tmp_path, case, impl):
    fn = splitter.split_data if impl == "native" else ref_split.split_data
    if impl == "oracle":
        os.makedirs(str(tmp_path) + ("/reverse_item-user" if case["reverse_user_item_data"] else ""), exist_ok=True)
    out = _run(fn, case, tmp_path)
    assert case["files"], "every case has at least the mymedialite CSVs"
    for name in case["files"]:
        assert filecmp.cmp(os.path.join(out, name), os.path.join(_golden_dir(case), name), shallow=False), name
    stream_after = np.random.random_sample()
    np.random.seed(case["seed"])
    np.random.permutation(sum(1 for _ in open(os.path.join(SPLIT, case["name"], "ratings.csv"))) - 1)
    assert stream_after == np.random.random_sample()          # one permutation drawn, nothing else


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_native_splitter_equals_oracle_everywhere(tmp_path, case):
    """Including the files the reference dies in and the unique-id lists."""
    a, b = tmp_path / "native", tmp_path / "oracle"
    os.makedirs(str(b) + ("/reverse_item-user" if case["reverse_user_item_data"] else ""))
    out_a = _run(splitter.split_data, case, a, save_users_and_items=True)
    out_b = _run(ref_split.split_data, case, b, save_users_and_items=True)
    names = sorted(os.listdir(out_b))
    assert sorted(os.listdir(out_a)) == names and len(names) == 7
    for name in names:
        assert filecmp.cmp(os.path.join(out_a, name), os.path.join(out_b, name), shallow=False), name
        if name.endswith(".json"):
            with open(os.path.join(out_a, name)) as f:
                json.load(f)


_field = st.one_of(st.integers(-50, 5000).map(str), st.floats(-1e6, 1e6, allow_nan=False).map(repr),
                   st.sampled_from(["0.5", "3", "1e-7", "2.50", "1E3", "12345678901234567890", "-0.0", ""]))
_ident = st.one_of(st.integers(0, 30).map(str), st.sampled_from(["7.0", "7.5", "u1", "a,b", 'q"r', "é", "x y", "007"]))


@settings(max_examples=100, deadline=None)
@given(rows=st.lists(st.tuples(_ident, _ident, _field, st.integers(0, 10 ** 9).map(str)), min_size=1, max_size=40),
       schema=st.sampled_from(["amazon", "netflix", "movielens", "yelp"]), ts=st.booleans(), rev=st.booleans(),
       seed=st.integers(0, 10 ** 6), fr=st.sampled_from([(.8, .1, .1), (.5, .25, .25), (1.0, 0.0, 0.0), (.34, .33, .33)]))
def test_native_splitter_equals_oracle_on_random_csvs(tmp_path_factory, rows, schema, ts, rev, seed, fr):
    d = str(tmp_path_factory.mktemp("csv")) + "/"
    ncol = 3 if schema == "netflix" else 4
    ts = ts and ncol == 4

    def q(s):
        return '"' + s.replace('"', '""') + '"' if any(c in s for c in ',"\n') else s
    with open(d + "r.csv", "w", encoding="utf-8") as f:
        f.write(",".join(["a", "b", "c", "d"][:ncol]) + "\n")
        for r in rows:
            f.write(",".join(q(x) for x in r[:ncol]) + "\n")
    os.makedirs(d + "n"), os.makedirs(d + "o" + ("/reverse_item-user" if rev else ""))
    results = []
    for fn, sub in ((splitter.split_data, "n/"), (ref_split.split_data, "o/")):
        np.random.seed(seed)
        try:
            results.append(fn(d + "r.csv", d + sub, schema, fr, True, ts, True, rev))
        except (ValueError, _lib.OcfError) as e:          # e.g. int('u1') under the movielens schema: both must refuse
            results.append(type(e))
    if isinstance(results[1], type):
        assert isinstance(results[0], type)
    else:
        for name in sorted(os.listdir(results[1])):
            assert filecmp.cmp(os.path.join(results[0], name), os.path.join(results[1], name), shallow=False), name
    shutil.rmtree(d)


def test_float_formatting_matches_python_repr_over_the_double_range(tmp_path):
    """json.dump / to_csv print floats with float.__repr__: shortest round-trip digits, fixed notation for
    1e-4 <= |x| < 1e16, else exponent form. 4 000 ratings spread over every binade (+ the notation boundaries,
    subnormals, the largest double) go through the native splitter and the Python oracle."""
    import random
    import struct
    rnd = random.Random(1)
    vals = [1e16, 9999999999999998.0, 1e15, 0.0001, 0.00009999, 5e-324, 1.7976931348623157e308, 2.2250738585072014e-308,
            1e22, 1e23, 0.1, 0.30000000000000004, 1 / 3, 100.0, 1e-7, 2.675, 1e21, -0.0, -1e-320, 12345678901234567.0]
    while len(vals) < 4000:
        v = struct.unpack("<d", struct.pack("<Q", rnd.getrandbits(64)))[0]
        if v == v and abs(v) != float("inf"):
            vals.append(v)
            vals.append(round(rnd.uniform(-1000, 1000), rnd.randint(0, 6)))
    d = str(tmp_path) + "/"
    with open(d + "r.csv", "w") as f:
        f.write("u,i,r,t\n")
        for k, v in enumerate(vals):
            f.write("%d,%s,%s,%d\n" % (k % 50, repr(float(k % 97) + 0.25 * (k % 3)), repr(v), k))
    os.makedirs(d + "n"), os.makedirs(d + "o")
    for fn, sub in ((splitter.split_data, "n/"), (ref_split.split_data, "o/")):
        np.random.seed(0)
        fn(d + "r.csv", d + sub, "amazon", (.8, .1, .1), True, True, True, False)
    for name in sorted(os.listdir(d + "o")):
        assert filecmp.cmp(d + "n/" + name, d + "o/" + name, shallow=False), name


# ---- ingest ------------------------------------------------------------------------------------------

def _python_ingest(path, col_of, paired, n_cols):
    with open(path) as f:
        obj = json.load(f)
    if not paired:
        return list(obj.keys()), _csr_from_lists(list(obj.values()), col_of, n_cols)
    ins, tgs = obj
    keys = list(tgs.keys())
    return keys, _csr_from_lists([ins[k] for k in keys], col_of, n_cols), np.array([ins[k] is None for k in keys]), \
        _csr_from_lists([tgs[k] for k in keys], col_of, n_cols)


def _same_csr(a, b):
    assert a.n_rows == b.n_rows and np.array_equal(a.rowptr, b.rowptr)
    assert np.array_equal(a.col, b.col) and np.array_equal(a.val, b.val, equal_nan=True) and a.val.dtype == b.val.dtype == np.float32


def _check_file(path, vocab_path, paired):
    with open(vocab_path) as f:
        ids = json.load(f)
    col_of = {x: i for i, x in enumerate(ids)}
    got = ingest.load_ratings(path, ingest.Vocab(vocab_path), paired)
    want = _python_ingest(path, col_of, paired, len(ids))
    assert got[0] == want[0]
    _same_csr(got[1], want[1])
    if paired:
        assert np.array_equal(got[2], want[2])
        _same_csr(got[3], want[3])
    return got


@pytest.mark.parametrize("name", ["ml", "ml_rev", "amazon"])
def test_native_ingest_equals_json_load_on_reference_files(tmp_path, name):
    case = [c for c in CASES if c["name"] == name][0]
    d = _golden_dir(case)
    with open(os.path.join(d, "ratingsByUser_dicts_train.json")) as f:
        train = json.load(f)
    ids = []
    for lst in train.values():
        ids += [p[0] for p in lst]
    for which in ("valid", "test"):
        with open(os.path.join(d, "ratingsByUser_dicts_%s.json" % which)) as f:
            for half in json.load(f):
                for lst in half.values():
                    ids += [p[0] for p in (lst or [])]
    ids = list(dict.fromkeys(int(x) if isinstance(x, float) else x for x in ids))     # 153.0 in the rows, 153 in the list
    vocab = str(tmp_path / "unique.json")
    with open(vocab, "w") as f:
        json.dump(ids, f)
    keys, csr = _check_file(os.path.join(d, "ratingsByUser_dicts_train.json"), vocab, False)
    assert csr.nnz > 150 and len(keys) > 20
    for which in ("valid", "test"):
        got = _check_file(os.path.join(d, "ratingsByUser_dicts_%s.json" % which), vocab, True)
        assert got[3].nnz > 10


def test_native_ingest_json_corner_cases(tmp_path):
    vocab = str(tmp_path / "v.json")
    with open(vocab, "w") as f:
        f.write(' [ 153, 136.0, "x\\u00e9\\ud83d\\ude00", 7.5, 1e3, -0.0, "153", 153 , 9007199254740993]\n')
    text = ('{ "a" : [[153.0, 4.5], ["x\\u00e9\\ud83d\\ude00", 3], [7.5, 1e0], [1000, 2.5E-1]],\n "b":[],\t"c":null,'
            ' "k\\"\\\\\\n\\u0041": [[136, 2.5, "extra", {"x": [1, 2]}], [0, -1.5], ["153", 0.1]],'
            ' "a": [[153, 1.0], [1.53e2, 2.0], [9007199254740993, 0.30000000000000004]], "b": [[-0, NaN]] }')
    path = str(tmp_path / "t.json")
    with open(path, "w") as f:
        f.write(text)
    keys, csr = _check_file(path, vocab, False)
    assert keys == ["a", "b", "c", 'k"\\\nA']                    # repeated keys keep their first place, last value
    assert csr.row(0)[0].tolist() == [7, 7, 8] and csr.row(1)[0].tolist() == [5]
    with open(path, "w") as f:
        f.write("[" + text + ", " + '{"c": [[153, 5]], "zz": [], "a": [[136, 1]]}' + "]")
    with pytest.raises(_lib.OcfError, match="KeyError.*zz"):
        ingest.load_ratings(path, ingest.Vocab(vocab), True)      # a target row the input dict does not hold
    with open(path, "w") as f:
        f.write("[" + text + ", " + '{"c": [[153, 5]], "b": [], "a": [[136, 1]]}' + "]")
    keys, ins, none, tgs = _check_file(path, vocab, True)
    assert keys == ["c", "b", "a"] and none.tolist() == [True, False, False]


@pytest.mark.parametrize("text,message", [
    ('{"a": [[99, 1.0]]}', "KeyError: 99"),
    ('{"a": [["q", 1.0]]}', "KeyError: 'q'"),
    ('{"a": [[153, 1.0]', "expected"),
    ('{"a": [[153 1.0]]}', "expected ','"),
    ('{"a": [[153, "5"]]}', "number"),
    ('{"a": [[153, 1.0]]} x', "trailing"),
    ('[[{}, {}], {}]', "_withtimestamps_"),
    ('{"a": [[153, 1.0, ' + "[" * 100000 + ']]}', "nesting"),
])
def test_native_ingest_errors(tmp_path, text, message):
    vocab = str(tmp_path / "v.json")
    with open(vocab, "w") as f:
        json.dump([153, 136], f)
    path = str(tmp_path / "t.json")
    with open(path, "w") as f:
        f.write(text)
    with pytest.raises(_lib.OcfError, match=message):
        ingest.load_ratings(path, ingest.Vocab(vocab), text.startswith("["))
    with pytest.raises(_lib.OcfError, match="cannot open"):
        ingest.load_ratings(path + ".missing", ingest.Vocab(vocab), False)


_jid = st.one_of(st.integers(-5, 40), st.floats(-5, 40, allow_nan=False, width=32), st.sampled_from(["a", "é", "15", ""]))
_jval = st.one_of(st.integers(-5, 5), st.floats(allow_nan=False, allow_infinity=False), st.floats(0.5, 5.0).map(lambda x: round(x, 1)))


@settings(max_examples=100, deadline=None)
@given(ids=st.lists(_jid, min_size=1, max_size=30),
       rows=st.dictionaries(st.text(max_size=5), st.one_of(st.none(), st.lists(st.tuples(st.integers(0, 29), _jval), max_size=8)), max_size=8),
       indent=st.sampled_from([None, 0, 2]), ascii_=st.booleans())
def test_native_ingest_equals_json_load_on_random_files(tmp_path_factory, ids, rows, indent, ascii_):
    d = str(tmp_path_factory.mktemp("json")) + "/"
    with open(d + "v.json", "w") as f:
        json.dump(ids, f, ensure_ascii=ascii_)
    obj = {k: (None if l is None else [[ids[i % len(ids)], v] for i, v in l]) for k, l in rows.items()}
    targets = {k: (l or []) for k, l in obj.items()}
    with open(d + "s.json", "w") as f:
        json.dump({k: l or [] for k, l in obj.items()}, f, indent=indent, ensure_ascii=ascii_)
    with open(d + "p.json", "w") as f:
        json.dump([obj, targets], f, indent=indent, ensure_ascii=ascii_)
    _check_file(d + "s.json", d + "v.json", False)
    _check_file(d + "p.json", d + "v.json", True)
    shutil.rmtree(d)


# ---- reference splitter output -> product reader, against the reference reader on the same files --------

def test_reader_on_reference_files_yields_the_reference_batches():
    d = os.path.join(SPLIT, "ml") + "/"
    gold = np.load(os.path.join(SPLIT, "pipeline_batches.npz"))
    rd = data_reader(int(gold["n_items"]), int(gold["n_rows"]), d, nonsequentialusers=False, use_json=True,
                     eval_mode="fixed_split", useTimestamps=False, reverse_user_item_data=False, rng_on_device=False)
    for which, sparsity, aux, seed in (("train", [0.3, 0.8], "dropout", 41), ("valid", None, None, 42), ("test", None, "both", 43)):
        np.random.seed(seed)
        gen = rd.data_gen(8, sparsity, train_val_test=which, shuffle=True, auxilliary_mask_type=aux, aux_var_value=-1,
                          return_target_count=which != "train")
        n = int(gold["%s/n_batches" % which])
        assert n > 0
        for b in range(n):
            batch = next(gen)
            feed, targets = host_densify(batch)
            assert np.array_equal(targets, gold["%s/b%d/targets" % (which, b)])
            for k, arr in enumerate(feed):
                assert np.array_equal(arr, gold["%s/b%d/in%d" % (which, b, k)]), (which, b, k)
            if which != "train":
                assert batch.target_count == int(gold["%s/b%d/target_count" % (which, b)])
        assert next(gen) is None
    rd.close()


def test_reader_in_ablation_mode_on_files_yields_the_reference_batches():
    """`ratingsByUser_dict.json` + `unique_users_list.json` (rows follow the LIST's order, data_reader.py:30-44,124-128)."""
    d = os.path.join(SPLIT, "ml_ablation") + "/"
    gold = np.load(os.path.join(SPLIT, "pipeline_batches.npz"))
    rd = data_reader(int(gold["n_items"]), int(gold["ablation/n_users"]), d, nonsequentialusers=True, use_json=True,
                     eval_mode="ablation", useTimestamps=False, reverse_user_item_data=False, rng_on_device=False)
    np.random.seed(51)
    rd.split_for_validation([0.6, 0.2, 0.2])
    for which, sparsity, seed in (("train", [0.2, 0.7], 52), ("test", [0.5, 0.5], 53)):
        np.random.seed(seed)
        gen = rd.data_gen(4, sparsity, train_val_test=which, shuffle=True, auxilliary_mask_type="causal", aux_var_value=-1)
        n = int(gold["ablation/%s/n_batches" % which])
        assert n > 0
        for b in range(n):
            feed, targets = host_densify(next(gen))
            assert np.array_equal(targets, gold["ablation/%s/b%d/targets" % (which, b)])
            for k, arr in enumerate(feed):
                assert np.array_equal(arr, gold["ablation/%s/b%d/in%d" % (which, b, k)]), (which, b, k)
        assert next(gen) is None
    rd.close()


def test_reversed_data_is_found_under_either_file_name(tmp_path):
    """reverse_user_item_data=True: the reader asks for `ratingsByItem_*` (data_reader.py:46-49) but the reference's
    splitter writes `ratingsByUser_*` into reverse_item-user/ (TrainValidTestSplit.py:153-157); both are accepted, and
    the column vocabulary is `unique_users_list` (data_reader.py:20-23)."""
    case = [c for c in CASES if c["name"] == "ml_rev"][0]
    src = _golden_dir(case)
    with open(os.path.join(src, "ratingsByUser_dicts_train.json")) as f:
        train = json.load(f)
    cols = []
    for which in ("train", "valid", "test"):
        with open(os.path.join(src, "ratingsByUser_dicts_%s.json" % which)) as f:
            obj = json.load(f)
        for half in (obj if which != "train" else [obj]):
            for lst in half.values():
                cols += [int(p[0]) for p in (lst or [])]
    cols = list(dict.fromkeys(cols))
    readers = []
    for name in ("ratingsByUser", "ratingsByItem"):
        d = str(tmp_path / name) + "/"
        os.makedirs(d)
        for which in ("train", "valid", "test"):
            shutil.copy(os.path.join(src, "ratingsByUser_dicts_%s.json" % which), d + "%s_dicts_%s.json" % (name, which))
        with open(d + "unique_users_list.json", "w") as f:
            json.dump(cols, f)
        readers.append(data_reader(len(cols), len(train), d, use_json=True, eval_mode="fixed_split",
                                   reverse_user_item_data=True, rng_on_device=False))
    a, b = readers
    assert a.train_set == b.train_set == list(train.keys()) and a.val_set == b.val_set
    for which in ("train",):
        _same_csr(a.store(which).csr, b.store(which).csr)
    assert a.store("train").csr.nnz == sum(len(l) for l in train.values())
    for r in readers:
        r.close()


def test_splitter_rejects_bad_arguments(tmp_path):
    csv = os.path.join(SPLIT, "ml", "ratings.csv")
    for bad in ((1.2, 0.1, 0.1), (-0.1, 0.5, 0.5), (0.7, 0.7, 0.1), (float("nan"), 0.1, 0.1)):
        with pytest.raises(_lib.OcfError, match="fraction"):
            splitter.split_data(csv, str(tmp_path) + "/", "movielens", bad, include_timestamps=False)
    with pytest.raises(_lib.OcfError, match="header has 4 columns"):
        splitter.split_data(csv, str(tmp_path) + "/", "netflix", include_timestamps=False)
    with pytest.raises(_lib.OcfError, match="timestamp"):
        splitter.split_data(os.path.join(SPLIT, "netflix_int", "ratings.csv"), str(tmp_path) + "/", "netflix", include_timestamps=True)
    with pytest.raises(ValueError):
        splitter.split_data(csv, str(tmp_path) + "/", "lastfm")
    # ids at the ends of the int64 range do not upset the id -> column window
    v = tmp_path / "v.json"
    v.write_text("[9223372036854775807, -9223372036854775807, 5]")
    t = tmp_path / "t.json"
    t.write_text('{"a": [[-9223372036854775807, 1.5], [5, 2], [9223372036854775807, 3]]}')
    keys, csr = ingest.load_ratings(str(t), ingest.Vocab(str(v)), False)
    assert csr.col.tolist() == [1, 2, 0]


def test_pickle_files_give_the_same_stores_as_json_files(tmp_path):
    """`use_json=False`: the reference's other on-disk format (`.p` pickles of the same objects, data_reader.py:88-91).
    Tuples inside pickles (the splitter's in-memory `(item, rating)` pairs) are accepted like lists."""
    import pickle
    src = os.path.join(SPLIT, "ml") + "/"
    d = str(tmp_path) + "/"
    for name in ("unique_items_list", "ratingsByUser_dicts_train", "ratingsByUser_dicts_valid", "ratingsByUser_dicts_test"):
        with open(src + name + ".json") as f:
            obj = json.load(f)
        if name.endswith("train"):
            obj = {k: [tuple(p) for p in v] for k, v in obj.items()}
        elif "dicts" in name:
            obj = tuple({k: (None if v is None else [tuple(p) for p in v]) for k, v in half.items()} for half in obj)
        with open(d + name + ".p", "wb") as f:
            pickle.dump(obj, f, protocol=2)
    gold = np.load(os.path.join(SPLIT, "pipeline_batches.npz"))
    a = data_reader(int(gold["n_items"]), int(gold["n_rows"]), src, use_json=True, eval_mode="fixed_split", rng_on_device=False)
    b = data_reader(int(gold["n_items"]), int(gold["n_rows"]), d, use_json=False, eval_mode="fixed_split", rng_on_device=False)
    assert a.train_set == b.train_set and a.val_set == b.val_set and a.test_set == b.test_set
    _same_csr(a.store("train").csr, b.store("train").csr)
    for which in ("valid", "test"):
        _same_csr(a.store(which).in_store.csr, b.store(which).in_store.csr)
        _same_csr(a.store(which).tgt_store.csr, b.store(which).tgt_store.csr)
    a.close(), b.close()


def _readers_agree(csv, schema, rev, seed, tmp, fractions=(.8, .1, .1)):
    """split_in_memory == split_data (files) + native ingest, store by store."""
    np.random.seed(seed)
    try:
        ls = splitter.split_in_memory(csv, schema, fractions, rev)
    except _lib.OcfError as e:
        ls = e
    np.random.seed(seed)
    try:
        d = splitter.split_data(csv, str(tmp) + "/", schema, fractions, True, False, True, rev)
        vocab = ingest.Vocab(d + "unique_items_list.json")
        files = (ingest.load_ratings(d + "ratingsByUser_dicts_train.json", vocab, False, vocab.size),
                 ingest.load_ratings(d + "ratingsByUser_dicts_valid.json", vocab, True, vocab.size),
                 ingest.load_ratings(d + "ratingsByUser_dicts_test.json", vocab, True, vocab.size))
    except _lib.OcfError as e:
        files = e
    if isinstance(files, Exception) or isinstance(ls, Exception):
        assert isinstance(files, Exception) and isinstance(ls, Exception), (files, ls)
        return None
    with open(d + "unique_items_list.json") as f:
        assert ls.unique_items == json.load(f) and ls.n_cols == vocab.size
    assert ls.train[0] == files[0][0]
    _same_csr(ls.train[1], files[0][1])
    for got, want in ((ls.valid, files[1]), (ls.test, files[2])):
        assert got[0] == want[0] and np.array_equal(got[2], want[2])
        _same_csr(got[1], want[1])
        _same_csr(got[3], want[3])
    return ls


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_split_in_memory_equals_the_path_through_files(tmp_path, case):
    csv = os.path.join(SPLIT, case["name"], "ratings.csv")
    ls = _readers_agree(csv, case["schema_type"], case["reverse_user_item_data"], case["seed"], tmp_path)
    assert ls is not None and ls.train[1].nnz > 100
    rd = data_reader(ls.n_cols + 3, len(ls.train[0]), "", eval_mode="fixed_split", data=ls, rng_on_device=False)
    assert rd.store("train").n_cols == ls.n_cols + 3 and rd.train_set == ls.train[0]       # arrays are num_items wide
    np.random.seed(1)
    batch = next(rd.data_gen(4, [0.5, 0.5], "train", True, "dropout", -1))
    feed, targets = host_densify(batch)
    assert targets.shape == (4, ls.n_cols + 3) and np.count_nonzero(feed[0]) + np.count_nonzero(targets) > 0
    rd.close()


@settings(max_examples=60, deadline=None)
@given(rows=st.lists(st.tuples(_ident, _ident, _field, st.integers(0, 10 ** 9).map(str)), min_size=1, max_size=40),
       schema=st.sampled_from(["amazon", "netflix", "movielens"]), rev=st.booleans(), seed=st.integers(0, 10 ** 6),
       fr=st.sampled_from([(.8, .1, .1), (.5, .25, .25), (1.0, 0.0, 0.0)]))
def test_split_in_memory_equals_the_path_through_files_on_random_csvs(tmp_path_factory, rows, schema, rev, seed, fr):
    d = str(tmp_path_factory.mktemp("mem")) + "/"
    ncol = 3 if schema == "netflix" else 4

    def q(s):
        return '"' + s.replace('"', '""') + '"' if any(c in s for c in ',"\n') else s
    with open(d + "r.csv", "w", encoding="utf-8") as f:
        f.write(",".join(["a", "b", "c", "d"][:ncol]) + "\n")
        for r in rows:
            f.write(",".join(q(x) for x in r[:ncol]) + "\n")
    os.makedirs(d + "out")
    _readers_agree(d + "r.csv", schema, rev, seed, d + "out", fr)
    shutil.rmtree(d)


def test_native_unique_lists_match_the_pipeline_fixture(tmp_path):
    case = [c for c in CASES if c["name"] == "ml"][0]
    out = _run(splitter.split_data, case, tmp_path, save_users_and_items=True)
    for name in ("unique_items_list.json", "unique_users_list.json"):
        assert filecmp.cmp(os.path.join(out, name), os.path.join(SPLIT, "ml", name), shallow=False)
