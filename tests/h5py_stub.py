"""A stand-in for the small part of the h5py API `checkpoint.py` uses (File / groups / datasets / attrs, nested
names like "dense_1/kernel:0"), for images without h5py: the tree lives in memory and is pickled behind the HDF5
magic bytes, so the format sniffing of `checkpoint.file_format` sees an "HDF5" file. Test infrastructure only."""
import pickle

import numpy as np

MAGIC = b"\x89HDF\r\n\x1a\n"


class Dataset(object):
    def __init__(self, shape, dtype):
        self.value = np.zeros(shape, dtype=dtype)
        self.shape, self.dtype = self.value.shape, self.value.dtype

    def __setitem__(self, key, val):
        self.value[key] = val

    def __getitem__(self, key):
        return self.value[key]

    def __array__(self, dtype=None, copy=None):
        return self.value if dtype is None else self.value.astype(dtype)


class Group(object):
    def __init__(self):
        self.attrs = {}
        self.children = {}

    def _walk(self, name, create):
        node = self
        parts = [p for p in name.split("/") if p]
        for p in parts[:-1]:
            if p not in node.children:
                if not create:
                    raise KeyError(name)
                node.children[p] = Group()
            node = node.children[p]
        return node, parts[-1]

    def create_group(self, name):
        node, leaf = self._walk(name, True)
        node.children[leaf] = Group()
        return node.children[leaf]

    def create_dataset(self, name, shape, dtype=np.float32):
        node, leaf = self._walk(name, True)
        node.children[leaf] = Dataset(shape, dtype)
        return node.children[leaf]

    def __getitem__(self, name):
        node, leaf = self._walk(name, False)
        return node.children[leaf]

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __delitem__(self, name):
        node, leaf = self._walk(name, False)
        del node.children[leaf]

    def keys(self):
        return self.children.keys()


class File(Group):
    def __init__(self, path, mode="r"):
        Group.__init__(self)
        self.path, self.mode = path, mode
        if mode in ("r", "r+", "a"):
            with open(path, "rb") as f:
                assert f.read(len(MAGIC)) == MAGIC
                tree = pickle.load(f)
            self.attrs, self.children = tree.attrs, tree.children

    def close(self):
        if self.mode != "r":
            tree = Group()
            tree.attrs, tree.children = self.attrs, self.children
            with open(self.path, "wb") as f:
                f.write(MAGIC)
                pickle.dump(tree, f)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
