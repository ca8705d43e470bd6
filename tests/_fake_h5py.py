"""TEST INFRASTRUCTURE: an in-memory stand-in for the few h5py calls the
reference's HDF5FileHandler makes (h5py is not in this image).  Files live in
a module-level dict keyed by path, so that a dump and a later load see the
same tree; `paths()` lists every group / dataset path for layout checks."""
import numpy as np

_FILES = {}


class Dataset:
    def __init__(self, data):
        self.data = np.array(data)

    def __getitem__(self, key):
        return self.data[key]

    @property
    def shape(self):
        return self.data.shape


class _Attrs(dict):
    pass


class Group:
    def __init__(self):
        self.children = {}
        self.attrs = _Attrs()

    def _walk(self, path, create):
        node = self
        for part in [p for p in str(path).split('/') if p]:
            if part not in node.children:
                if not create:
                    return None
                node.children[part] = Group()
            node = node.children[part]
        return node

    def require_group(self, name):
        return self._walk(name, True)

    create_group = require_group

    def create_dataset(self, name, data=None, **_):
        if name in self.children:
            raise ValueError(f'Unable to create dataset (name already exists): {name}')
        self.children[name] = Dataset(data)
        return self.children[name]

    def get(self, name, default=None):
        node = self._walk(name, False)
        return default if node is None else node

    def __getitem__(self, name):
        node = self._walk(name, False)
        if node is None:
            raise KeyError(name)
        return node

    def __contains__(self, name):
        return self._walk(name, False) is not None

    def __delitem__(self, name):
        del self.children[name]

    def keys(self):
        return self.children.keys()

    def paths(self, prefix=''):
        out = []
        for k, v in self.children.items():
            out.append(f'{prefix}/{k}')
            if isinstance(v, Group):
                out += v.paths(f'{prefix}/{k}')
        return out


class File(Group):
    def __new__(cls, path, mode='r'):
        key = str(path)
        if key not in _FILES:
            if mode == 'r':
                raise OSError(f'Unable to open file {key}')
            inst = super().__new__(cls)
            Group.__init__(inst)
            _FILES[key] = inst
        return _FILES[key]

    def __init__(self, path, mode='r'):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def flush(self):
        pass

    def close(self):
        pass
