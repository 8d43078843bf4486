"""Minimal stand-in for mmcv.Config / Registry so the reference's python configs
(configs/multi/*.py, `_base_`, `_delete_`, `{{_base_.x}}`, custom_imports) load
UNMODIFIED without mmcv (SURVEY D.6, 8b 'Config schema').

Reference: tools/train.py:119-128 (Config.fromfile, merge_from_dict),
mtl/data/build.py:31-40 (load_data_cfg).
"""
import copy
import os
import re

DELETE_KEY = '_delete_'
BASE_KEY = '_base_'


class ConfigDict(dict):
    """dict with attribute access (mmcv ConfigDict / addict semantics)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError("'ConfigDict' object has no attribute '%s'" % name)

    def __setattr__(self, name, value):
        self[name] = _wrap(value)

    def __delattr__(self, name):
        del self[name]

    def copy(self):
        return ConfigDict(dict.copy(self))

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _wrap(v):
    if isinstance(v, dict) and not isinstance(v, ConfigDict):
        return ConfigDict({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, list):
        return [_wrap(x) for x in v]
    if isinstance(v, tuple):
        return tuple(_wrap(x) for x in v)
    return v


def _merge(a, b):
    """merge child dict a into base dict b (mmcv Config._merge_a_into_b)."""
    b = dict(b)
    for k, v in a.items():
        if isinstance(v, dict) and k in b and isinstance(b[k], dict) and not v.get(DELETE_KEY, False):
            b[k] = _merge(v, b[k])
        else:
            if isinstance(v, dict):
                v = {kk: vv for kk, vv in v.items() if kk != DELETE_KEY}
            b[k] = v
    return b


_BASE_VAR = re.compile(r'\{\{\s*_base_\.([\w\.]+)\s*\}\}')


def _file2dict(filename):
    filename = os.path.abspath(filename)
    if not os.path.isfile(filename):
        raise FileNotFoundError('config file %s does not exist' % filename)
    text = open(filename).read()
    # textual substitution of {{_base_.name}} by a placeholder resolved after the bases are loaded
    placeholders = {}

    def repl(m):
        key = '_BASEVAR_%d_' % len(placeholders)
        placeholders[key] = m.group(1)
        return key

    text = _BASE_VAR.sub(repl, text)
    # first pass only to discover _base_
    m = re.search(r'^_base_\s*=\s*(.+?)$', text, flags=re.M | re.S)
    base_cfg = {}
    scope = {}
    if m:
        probe = {}
        # _base_ may span several lines (list); exec only that statement safely
        stmt = _extract_statement(text, m.start())
        exec(stmt, {}, probe)
        bases = probe[BASE_KEY]
        bases = [bases] if isinstance(bases, str) else list(bases)
        for b in bases:
            sub = _file2dict(os.path.join(os.path.dirname(filename), b))
            dup = set(base_cfg) & set(sub)
            if dup:
                raise KeyError('Duplicate key is not allowed among bases: %s' % dup)
            base_cfg.update(sub)
    for key, path in placeholders.items():
        v = base_cfg
        for part in path.split('.'):
            v = v[part]
        scope[key] = copy.deepcopy(v)
    g = {'__file__': filename}
    g.update(scope)
    exec(compile(text, filename, 'exec'), g)
    cfg = {k: v for k, v in g.items()
           if not k.startswith('__') and k not in scope and not callable(v) and not isinstance(v, type(os))}
    cfg.pop(BASE_KEY, None)
    return _merge(cfg, base_cfg)


def _extract_statement(text, start):
    """return the full (possibly multi-line) assignment statement starting at `start`."""
    depth = 0
    i = start
    n = len(text)
    while i < n:
        c = text[i]
        if c in '([{':
            depth += 1
        elif c in ')]}':
            depth -= 1
        elif c == '\n' and depth == 0:
            # continuation by backslash or implicit string concatenation on the next line
            j = i - 1
            while j >= 0 and text[j] in ' \t':
                j -= 1
            if text[j] != '\\':
                break
        i += 1
    return text[start:i]


class Config:
    def __init__(self, cfg_dict=None, filename=None):
        object.__setattr__(self, '_cfg_dict', _wrap(cfg_dict or {}))
        object.__setattr__(self, 'filename', filename)

    @staticmethod
    def fromfile(filename):
        cfg = Config(_file2dict(filename), filename)
        ci = cfg._cfg_dict.get('custom_imports')
        if ci:
            import importlib
            mods = ci['imports']
            for mname in ([mods] if isinstance(mods, str) else mods):
                try:
                    importlib.import_module(_IMPORT_ALIASES.get(mname, mname))
                except ImportError:
                    if not ci.get('allow_failed_imports', False):
                        raise
        return cfg

    def merge_from_dict(self, options):
        d = {}
        for full_key, v in options.items():
            cur = d
            keys = full_key.split('.')
            for k in keys[:-1]:
                cur = cur.setdefault(k, {})
            cur[keys[-1]] = v
        object.__setattr__(self, '_cfg_dict', _wrap(_merge(d, self._cfg_dict)))

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __setattr__(self, name, value):
        self._cfg_dict[name] = _wrap(value)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setitem__(self, name, value):
        self._cfg_dict[name] = _wrap(value)

    def __contains__(self, name):
        return name in self._cfg_dict

    def get(self, key, default=None):
        return self._cfg_dict.get(key, default)

    def copy(self):
        return Config(copy.deepcopy(self._cfg_dict), self.filename)


# the reference registers its classes by importing `models.multi`; here the same
# registry names are provided by this package.
_IMPORT_ALIASES = {'models.multi': 'rscotr_b200.models', 'models.det': 'rscotr_b200.models',
                   'models.seg': 'rscotr_b200.models'}


class Registry:
    def __init__(self, name):
        self.name = name
        self._map = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            key = name or cls.__name__
            if key in self._map and not force:
                raise KeyError('%s is already registered in %s' % (key, self.name))
            self._map[key] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self._map.get(key)

    def build(self, cfg, **default_args):
        return build_from_cfg(cfg, self, default_args)

    def __contains__(self, key):
        return key in self._map


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict):
        raise TypeError('cfg must be a dict, but got %s' % type(cfg))
    if 'type' not in cfg:
        raise KeyError('`cfg` must contain the key "type", but got %s' % cfg)
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    t = args.pop('type')
    if isinstance(t, str):
        cls = registry.get(t)
        if cls is None:
            raise KeyError('%s is not in the %s registry' % (t, registry.name))
    else:
        cls = t
    return cls(**args)


MODELS = Registry('models')
