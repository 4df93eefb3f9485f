"""
jlrt -- runtime for Python code produced by oracle/jl2py.py.  TEST INFRASTRUCTURE ONLY.

Reproduces the Julia 0.6 semantics the reference WENO5 files rely on:
  * arrays are column-major (numpy order="F"), indices are 1-based, `a[i,:]` copies, scalar indices drop
    the dimension, a single index into an N-d array is a linear (column-major) index;
  * under `@inbounds` an index that leaves its own dimension but stays inside the allocation reads the
    linear memory location (the reference does this on the high side of every sweep); a read
    outside the allocation returns NaN here and is counted in `OOB["outside"]` (Julia would return
    whatever follows the array on the heap);
  * `x^2` is `x*x` (Julia lowers literal powers 2 and 3 to multiplications), `x^y` with a float y is libm pow;
  * `max`/`min` propagate NaN; numbers can be "indexed" (`dS = 1; dS[i,j]`);
  * uninitialised `Array{Float64}(n,m)` is zero-filled.
"""
import math

import numpy as np

np.seterr(all="ignore")

PI = math.pi
INF = math.inf
NAN = math.nan
Float64 = float
String = str
Bool = bool
Int16 = np.int16
Int32 = np.int32


class _Colon:
    def __repr__(self):
        return "COLON"


COLON = _Colon()


class _End:
    """`end` inside an index expression; supports end±k, end/2 is not needed."""
    __slots__ = ("off",)

    def __init__(self, off=0):
        self.off = off

    def __add__(self, k):
        return _End(self.off + k)

    __radd__ = __add__

    def __sub__(self, k):
        return _End(self.off - k)

    def resolve(self, n):
        return n + self.off


END = _End()


class jrange:
    """a:b or a:s:b (inclusive, 1-based when used as an index)."""
    __slots__ = ("a", "b", "s")

    def __init__(self, a, b, s=1):
        self.a, self.b, self.s = a, b, s

    def __iter__(self):
        a, b, s = self.a, self.b, self.s
        if isinstance(a, int) and isinstance(b, int) and isinstance(s, int):
            return iter(range(a, b + (1 if s > 0 else -1), s))
        n = int(math.floor((b - a) / s + 1e-12)) + 1
        return iter([a + i * s for i in range(max(n, 0))])

    def __len__(self):
        return len(list(iter(self)))

    def resolve(self, n):
        a = self.a.resolve(n) if isinstance(self.a, _End) else self.a
        b = self.b.resolve(n) if isinstance(self.b, _End) else self.b
        return jrange(a, b, self.s)


def jiter(x):
    if isinstance(x, np.ndarray):
        return iter(x.reshape(-1, order="F"))
    return iter(x)


OOB = {"wrapped": 0, "outside": 0}


def _lin(a):
    """1-D view of the column-major memory of a."""
    if a.ndim <= 1:
        return a
    assert a.flags.f_contiguous, "jlrt arrays must be column-major"
    return a.reshape(-1, order="F")


def _conv_index(a, idx):
    """Translate a Julia index tuple into a numpy index; returns (numpy_index, all_scalar)."""
    nd = a.ndim
    if len(idx) == 1 and nd > 1:
        return None, False
    out = []
    scalar = True
    for d, ix in enumerate(idx):
        n = a.shape[d] if d < nd else 1
        if isinstance(ix, _End):
            ix = ix.resolve(n)
        if isinstance(ix, (int, np.integer)):
            out.append(int(ix) - 1)
        elif ix is COLON:
            out.append(slice(None))
            scalar = False
        elif isinstance(ix, jrange):
            r = ix.resolve(n)
            if not (1 <= r.a and r.b <= n) and r.b >= r.a:
                raise IndexError(f"range {r.a}:{r.b} outside dimension of length {n}")
            out.append(slice(r.a - 1, r.b, r.s) if r.s > 0 else slice(r.a - 1, (r.b - 2) if r.b >= 2 else None, r.s))
            scalar = False
        elif isinstance(ix, (list, np.ndarray)):
            arr = np.asarray(ix)
            if arr.dtype == bool:
                out.append(arr)
            else:
                out.append(arr.astype(np.int64) - 1)
            scalar = False
        elif isinstance(ix, float) and ix == int(ix):
            out.append(int(ix) - 1)
        else:
            raise IndexError(f"unsupported index {ix!r}")
    return out, scalar


def jget(a, idx):
    if not isinstance(a, np.ndarray):
        if isinstance(a, (tuple, list)):
            return a[idx[0] - 1]
        return a                                     # number indexed like an array (@inbounds)
    ni, scalar = _conv_index(a, idx)
    if ni is None:                                   # linear indexing
        ix = idx[0]
        flat = _lin(a)
        if isinstance(ix, _End):
            ix = ix.resolve(a.size)
        if isinstance(ix, (int, np.integer)):
            return flat[ix - 1]
        if ix is COLON:
            return flat.copy()
        if isinstance(ix, jrange):
            r = ix.resolve(a.size)
            return flat[r.a - 1:r.b:r.s].copy()
        ix = np.asarray(ix)
        if ix.dtype == bool:
            return flat[_lin(ix)]
        return flat[ix.astype(np.int64) - 1]
    if scalar:
        nd = a.ndim
        ni = ni[:nd] if len(ni) > nd and all(v == 0 for v in ni[nd:]) else ni
        ok = len(ni) == nd
        if ok:
            for v, n in zip(ni, a.shape):
                if v < 0 or v >= n:
                    ok = False
                    break
        if ok:
            return a[tuple(ni)]
        # @inbounds semantics: linear memory offset
        off, stride = 0, 1
        for d, v in enumerate(ni):
            off += v * stride
            stride *= a.shape[d] if d < nd else 1
        if 0 <= off < a.size:
            OOB["wrapped"] += 1
            return _lin(a)[off]
        OOB["outside"] += 1
        return np.float64(np.nan)
    # trailing singleton indices beyond ndim
    ni = ni[:a.ndim]
    n_adv = sum(isinstance(v, np.ndarray) for v in ni)
    if n_adv > 1:
        ni = np.ix_(*[np.arange(a.shape[d])[v] if not isinstance(v, int) else [v] for d, v in enumerate(ni)])
        r = a[ni]
        drop = tuple(d for d, v in enumerate(idx[:a.ndim]) if isinstance(v, (int, np.integer)))
        r = r.reshape([s for d, s in enumerate(r.shape) if d not in drop])
        return np.array(r, order="F")
    return np.array(a[tuple(ni)], order="F")


def jset(a, idx, v):
    ni, scalar = _conv_index(a, idx)
    if ni is None:
        ix = idx[0]
        flat = _lin(a)
        if isinstance(ix, _End):
            ix = ix.resolve(a.size)
        if isinstance(ix, (int, np.integer)):
            flat[ix - 1] = v
        elif ix is COLON:
            flat[:] = _lin(v) if isinstance(v, np.ndarray) else v
        elif isinstance(ix, jrange):
            r = ix.resolve(a.size)
            flat[r.a - 1:r.b:r.s] = _lin(v) if isinstance(v, np.ndarray) else v
        else:
            ix = np.asarray(ix)
            if ix.dtype == bool:
                flat[_lin(ix)] = v
            elif ix.size:
                flat[ix.astype(np.int64) - 1] = v
        return
    ni = ni[:a.ndim]
    if scalar:
        for val, n in zip(ni, a.shape):
            if val < 0 or val >= n:
                raise IndexError(f"write outside array: index {idx} shape {a.shape}")
    n_adv = sum(isinstance(x, np.ndarray) for x in ni)
    if n_adv > 1:
        raise IndexError("multiple vector indices in assignment not supported")
    if isinstance(v, np.ndarray) and not scalar:
        tgt = a[tuple(ni)]
        if v.shape != tgt.shape:
            if v.size != tgt.size:
                raise ValueError(f"shape mismatch in assignment: {v.shape} into {tgt.shape}")
            v = v.reshape(tgt.shape, order="F")
    a[tuple(ni)] = v


def jbroadcast_assign(lhs, value):
    lhs[...] = value


# ---------------------------------------------------------------- constructors
def _dims(args):
    dims = [a for a in args if not isinstance(a, type)]
    if len(dims) == 1 and isinstance(dims[0], tuple):
        dims = list(dims[0])
    return tuple(int(d) for d in dims)


def zeros(*args):
    return np.zeros(_dims(args), order="F")


def ones(*args):
    return np.ones(_dims(args), order="F")


def Array(*args):
    return np.zeros(_dims(args), order="F")


SharedArray = Array
Vector = Array
Matrix = Array


def fill(v, *dims):
    return np.full(_dims(dims), float(v), order="F")


def similar(a, *args):
    return np.zeros(a.shape, order="F")


def jcopy(a):
    return np.array(a, order="F") if isinstance(a, np.ndarray) else a


deepcopy = jcopy


def jvect(items):
    if any(isinstance(x, np.ndarray) for x in items):
        return np.concatenate([np.atleast_1d(x) for x in items])
    if all(isinstance(x, (int, np.integer)) and not isinstance(x, bool) for x in items):
        return np.array(items, dtype=np.int64)
    return np.array(items, dtype=np.float64)


def size(a, d=None):
    if d is None:
        return a.shape
    return a.shape[d - 1] if d <= a.ndim else 1


def jlength(a):
    return a.size if isinstance(a, np.ndarray) else len(a)


def ndims(a):
    return a.ndim


def transpose(a):
    if isinstance(a, np.ndarray):
        return np.array(a.T, order="F")
    return a


def permutedims(a, perm):
    return np.array(np.transpose(a, [p - 1 for p in perm]), order="F")


def reshape(a, *dims):
    return np.array(a.reshape(_dims(dims), order="F"), order="F")


def find(mask):
    if not isinstance(mask, np.ndarray):
        return np.array([1], dtype=np.int64) if mask else np.array([], dtype=np.int64)
    return np.flatnonzero(_lin(mask)) + 1


def linspace(a, b, n):
    return np.linspace(a, b, int(n))


def collect(x):
    return jvect(list(x))


# ---------------------------------------------------------------- arithmetic
def jpow(x, y):
    if isinstance(y, (int, np.integer)) and not isinstance(y, bool):
        if y == 2:
            return x * x
        if y == 3:
            return x * x * x
        if y >= 0:                                   # Base.power_by_squaring
            if y == 0:
                return x * 0 + 1
            t = (y & -y).bit_length()                # trailing zeros + 1
            y >>= t
            while t > 1:
                x = x * x
                t -= 1
            r = x
            while y > 0:
                t = (y & -y).bit_length()
                y >>= t
                while t > 0:
                    x = x * x
                    t -= 1
                r = r * x
            return r
        if isinstance(x, (int, np.integer)):
            raise ValueError("integer to negative integer power")
        return jpow(1.0 / x, -y)
    if isinstance(x, np.ndarray) or isinstance(y, np.ndarray):
        return np.power(x, y)
    return np.float64(math.pow(x, y)) if x >= 0 or float(y).is_integer() else np.float64(np.nan)


def sqrt(x):
    return np.sqrt(x)


def cbrt(x):
    return np.cbrt(x)


def exp(x):
    return np.exp(x)


def log(x):
    return np.log(x)


def log10(x):
    return np.log10(x)


def sin(x):
    return np.sin(x)


def cos(x):
    return np.cos(x)


def tan(x):
    return np.tan(x)


def atan2(y, x):
    return np.arctan2(y, x)


def jabs(x):
    return abs(x)


def abs2(x):
    return x * x


def sign(x):
    return np.sign(x)


def floor(x, *a):
    return np.floor(x)


def ceil(x, *a):
    return np.ceil(x)


def jround(x, *a):
    if isinstance(x, type) or x in (Int16, Int32):          # round(Int16, v): ties to even, then convert
        v = np.rint(a[0])
        return v.astype(x) if isinstance(v, np.ndarray) else int(v)
    return np.rint(x)


def jeq(a, b):
    if isinstance(a, (np.ndarray, list, tuple)) or isinstance(b, (np.ndarray, list, tuple)):
        a, b = np.asarray(a), np.asarray(b)
        return a.shape == b.shape and bool(np.all(a == b))
    return a == b


def jne(a, b):
    return not jeq(a, b)


def jInt(x):
    return int(x)


def jfloat(x):
    return np.float64(x) if not isinstance(x, np.ndarray) else x.astype(np.float64)


def isnan(x):
    return np.isnan(x)


def isinf(x):
    return np.isinf(x)


def isfinite(x):
    return np.isfinite(x)


def _minmax(args, npf, better):
    if any(isinstance(a, np.ndarray) for a in args):
        r = args[0]
        for a in args[1:]:
            r = npf(r, a)                            # np.maximum / np.minimum propagate NaN like Julia
        return r
    r = args[0]
    for a in args[1:]:
        if a != a:
            return a
        if r != r:
            return r
        if better(a, r):
            r = a
    return r


def jmax(*args):
    return _minmax(args, np.maximum, lambda a, r: a > r)


def jmin(*args):
    return _minmax(args, np.minimum, lambda a, r: a < r)


def maximum(a):
    return np.max(a)


def minimum(a):
    return np.min(a)


def jsum(a):
    return np.sum(a)


def mean(a):
    return np.mean(a)


def jany(a):
    return bool(np.any(a))


def jall(a):
    return bool(np.all(a))


def jnot(x):
    if isinstance(x, np.ndarray):
        return ~x
    return not x


def jrem(a, b):
    return math.fmod(a, b) if isinstance(a, float) or isinstance(b, float) else int(math.fmod(a, b))


def jdiv(a, b):
    return int(a / b)


def div(a, b):
    return int(a / b)


def mod(a, b):
    return a % b


def typeof(x):
    return type(x)


def isempty(a):
    return jlength(a) == 0


def error(*msg):
    raise RuntimeError(" ".join(str(m) for m in msg))
