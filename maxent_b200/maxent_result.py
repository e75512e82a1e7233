"""Result of a MaxEnt run: alpha-resolved arrays plus the analyzers' verdicts.

Public surface of python/maxent_result.py (``MaxEntResultData`` :157-685, ``MaxEntResult`` :688-1081):
fields ``alpha, v, chi2, S, Q, H, A, probability, G, G_orig, G_rec, data_variable, omega,
analyzer_results, run_times, run_time_total, matrix_structure, effective_matrix_structure, element_wise,
complex_elements, use_hermiticity, default_analyzer_name, zero_elements``; ``A_out / get_A_out``,
``default_analyzer / get_default_analyzer``, ``analyze``, ``include / exclude / include_only``,
``data`` (array-only twin), dict round trip for h5 / pickle.

The reference stores one bound cost-function OBJECT per alpha and rebuilds every array lazily from
them (python/maxent_result.py:720-791).  The fused path produces arrays, so this class stores one
*sweep record* (a dict of arrays for all alphas) per matrix element: ``add_sweep``.  ``add_result`` is
kept for callers that append alpha by alpha."""
import copy
from collections import OrderedDict
from datetime import datetime, timedelta
from itertools import product

import numpy as np

from .alpha_meshes import DataAlphaMesh
from .omega_meshes import DataOmegaMesh

_FIELDS = ['alpha', 'v', 'chi2', 'S', 'A', 'Q', 'omega', 'probability', 'analyzer_results', 'run_times',
           'run_time_total', 'matrix_structure', 'effective_matrix_structure', 'element_wise',
           'complex_elements', 'use_hermiticity', 'G', 'data_variable', 'G_rec', 'H',
           'default_analyzer_name', 'zero_elements', 'G_orig']


def saved(func):
    """Property computed once and kept in ``_saved`` (so that the array-only twin can serve it too)."""
    name = func.__name__

    def getter(self):
        if name not in self._all_fields and self._check_fieldnames:
            raise AttributeError('Field {} not available'.format(name))
        if name not in self._saved:
            self._saved[name] = func(self)
        return self._saved[name]
    getter.__name__ = name
    getter.__doc__ = func.__doc__
    return property(getter)


def recursive_map(seq, func):
    """Apply ``func`` to all non-list items of a nested list."""
    for item in seq:
        if isinstance(item, list):
            yield type(item)(recursive_map(item, func))
        else:
            yield func(item)


def recursive_dtype(seq):
    for item in seq:
        if isinstance(item, list):
            return recursive_dtype(item)
        return getattr(item, "dtype", type(item))


def _nested(structure, make):
    """Nested lists of shape ``structure`` whose leaves are ``make()``."""
    if not structure:
        return make()
    return [_nested(structure[1:], make) for _ in range(structure[0])]


def _td_to_dict(t):
    if isinstance(t, timedelta):
        return dict(days=t.days, seconds=t.seconds, microseconds=t.microseconds)
    if isinstance(t, float):
        return t
    return [_td_to_dict(x) for x in t]


def _td_from_dict(t):
    if isinstance(t, dict):
        return timedelta(**t)
    if isinstance(t, float):
        return t
    return [_td_from_dict(x) for x in t]


class MaxEntResultData(object):
    """Array-only result (what gets written to h5 / pickled)."""

    def __init__(self, matrix_structure=None, element_wise=True, complex_elements=False, use_hermiticity=True):
        self._all_fields = list(_FIELDS)
        self._matrix_structure = None if matrix_structure is None else tuple(matrix_structure)
        self._complex_elements = complex_elements
        self._element_wise = element_wise
        self._use_hermiticity = use_hermiticity
        self._check_fieldnames = True
        self._default_analyzer_name = None
        self._zero_elements = []
        self._saved = dict()

    def __getattr__(self, name):
        if name == "_saved":
            raise AttributeError(name)
        if name in self._saved:
            return self._saved[name]
        raise AttributeError("'{}' object has no attribute '{}'".format(type(self).__name__, name))

    def _get_element(self, array, matrix_element):
        if self.matrix_structure is None:
            assert matrix_element is None, "Cannot give matrix_element when matrix_structure is None"
            return array
        if not self._element_wise:
            assert matrix_element is None, "Cannot give matrix_element when element_wise is False"
            return array
        assert matrix_element is not None, "matrix_element must be given"
        out = array
        for i in matrix_element:
            out = out[i]
        return out

    # ---- which fields travel -------------------------------------------------------------------------
    def include_only(self, fields):
        self._all_fields = []
        self.include(fields)

    def include(self, fields):
        old, self._check_fieldnames = self._check_fieldnames, False
        try:
            for f in fields:
                if not hasattr(self, f):
                    raise AttributeError('Unknown field: {}'.format(f))
                if f not in self._all_fields:
                    self._all_fields.append(f)
        finally:
            self._check_fieldnames = old

    def exclude(self, fields):
        for f in fields:
            if not hasattr(self, f):
                raise AttributeError('Unknown field: {}'.format(f))
            if f in self._all_fields:
                self._all_fields.remove(f)

    # ---- analyzers' output ---------------------------------------------------------------------------
    def get_default_analyzer(self, analyzer=None):
        """AnalyzerResult of ``analyzer`` (default: ``default_analyzer_name``, else LineFitAnalyzer); for an
        element-wise matrix result an object array of them (None where not computed)."""
        name = analyzer or self.default_analyzer_name or 'LineFitAnalyzer'
        if self.matrix_structure is None or not self.element_wise:
            return self.analyzer_results[name]
        out = np.empty(self.effective_matrix_structure, dtype=object)
        for elem in product(*map(range, self.effective_matrix_structure)):
            try:
                out[elem] = self._get_element(self.analyzer_results, elem)[name]
            except KeyError:
                out[elem] = None
        return out

    default_analyzer = property(get_default_analyzer)

    def get_A_out(self, analyzer=None):
        """The one spectral function chosen by ``analyzer``.  Matrix results: M x N x n_omega, elements that
        were not computed are taken from the transposed element when ``use_hermiticity`` (conjugated for
        complex elements), elements below G_threshold are zero, everything else NaN
        (python/maxent_result.py:314-366)."""
        da = self.get_default_analyzer(analyzer)
        if self.matrix_structure is None or not self.element_wise:
            return da['A_out']
        shape = self.effective_matrix_structure
        A_out = np.full(shape + (len(self.omega),), np.nan)
        for elem in self.zero_elements:
            A_out[elem] = 0.0
        for elem in product(*map(range, shape)):
            src, sign = elem, 1.0
            if da[elem] is None and self.use_hermiticity:
                src = (elem[1], elem[0]) + tuple(elem[2:])
                if self.complex_elements and src[-1] == 1:
                    sign = -1.0
            if da[src] is not None and 'A_out' in da[src]:
                A_out[elem] = sign * da[src]['A_out']
        if self.complex_elements:
            return A_out[..., 0, :] + 1.0j * A_out[..., 1, :]
        return A_out

    A_out = property(get_A_out)

    # ---- plot data (x, y, options) as in the reference's @plot_function providers ------------------------
    def _matrix_opts(self, d, check_element_wise=True):
        out = OrderedDict()
        if self.matrix_structure is not None and ((not check_element_wise) or self.element_wise):
            out['n_m'], out['n_n'] = self.matrix_structure[0], self.matrix_structure[1]
            if self.complex_elements:
                out['n_c'] = 2
        out.update(d)
        return out

    def _curve(self, y, label, element):
        idx = slice(None) if element is None else element
        return (self.alpha, y[idx], self._matrix_opts(OrderedDict(label=label, x_label=r'$\alpha$', y_label=label,
                                                                  log_x=True, log_y=True)))

    def plot_chi2(self, element=None, **kwargs):
        return self._curve(self.chi2, r'$\chi^2$', element)

    def plot_S(self, element=None, **kwargs):
        return self._curve(self.S, r'$S$', element)

    def plot_Q(self, element=None, **kwargs):
        return self._curve(self.Q, r'$Q$', element)

    def plot_probability(self, element=None, **kwargs):
        if np.all(np.isnan(self.probability)):
            raise AttributeError('Probability is all NaN')
        idx = slice(None) if element is None else element
        p = self.probability[idx]
        return (self.alpha, np.exp(p - np.nanmax(p)),
                self._matrix_opts(OrderedDict(label='$p$', x_label=r'$\alpha$', y_label='$p$', log_x=True, log_y=False)))

    def plot_A(self, element=None, alpha_index=0, **kwargs):
        """(omega, A_alpha(omega), options) of one alpha (python/maxent_result.py:468-501)."""
        idx = slice(None) if element is None else element
        return (self.omega, self.A[idx][alpha_index],
                self._matrix_opts(OrderedDict(label=r'$A_{{\alpha_{}}}(\omega)$'.format(alpha_index), x_label=r'$\omega$',
                                              y_label=r'$A(\omega)$', log_x=False, log_y=False,
                                              n_alpha_index=len(self.alpha)), check_element_wise=False))

    def plot_G(self, element=None, **kwargs):
        """(data variable, original G, options) (python/maxent_result.py:503-539)."""
        idx = slice(None) if element is None else element
        return (self.data_variable[idx], self.G_orig[idx],
                self._matrix_opts(OrderedDict(label=r'$G(d)$', x_label=r'$d$', y_label=r'$G(d)$', log_x=False, log_y=False),
                                  check_element_wise=False))

    def plot_G_rec(self, element=None, alpha_index=0, plot_G=True, **kwargs):
        """List of curves: the original data (if ``plot_G``) and the reconstruction G_rec = K_delta A_alpha
        (python/maxent_result.py:541-582)."""
        ret = []
        if plot_G:
            ret.append(self.plot_G(element=element, **kwargs))
        idx = slice(None) if element is None else element
        ret.append((self.data_variable[idx], self.G_rec[idx][alpha_index],
                    self._matrix_opts(OrderedDict(label=r'$G_{rec}(d)$', x_label=r'$d$', y_label=r'$G(d)$', log_x=False,
                                                  log_y=False, plot_G=True, n_alpha_index=len(self.alpha)),
                                      check_element_wise=False)))
        return ret

    # ---- dict round trip (h5 / pickle; python/maxent_result.py:616-685) ---------------------------------
    def __reduce_to_dict__(self):
        out = dict(all_fields=self._all_fields)
        for key in self._all_fields:
            val = getattr(self, key)
            out[key] = 'None' if val is None else val
        if 'run_times' in out:
            out['run_times'] = [_td_to_dict(t) for t in out['run_times']]
        if 'run_time_total' in out:
            out['run_time_total'] = _td_to_dict(out['run_time_total'])
        return out

    @classmethod
    def __factory_from_dict__(cls, name, D):
        self = cls()
        D = dict(D)
        if 'run_times' in D:
            D['run_times'] = [_td_from_dict(t) for t in D['run_times']]
        if 'run_time_total' in D:
            D['run_time_total'] = _td_from_dict(D['run_time_total'])
        if 'omega' in D:
            D['omega'] = DataOmegaMesh(D['omega'])
        if 'alpha' in D:
            D['alpha'] = DataAlphaMesh(D['alpha'])
        if 'all_fields' in D:
            self._all_fields = D.pop('all_fields')

        def attach(x):
            if isinstance(x, dict):
                for val in x.values():
                    if not isinstance(val, str):
                        val.maxent_result = self
            else:
                for y in x:
                    attach(y)
        if 'analyzer_results' in D:
            attach(D['analyzer_results'])
        for key, val in D.items():
            self._saved[key] = None if (isinstance(val, str) and val == 'None') else val
        return self

    def __getstate__(self):
        return self.__dict__

    def __setstate__(self, d):
        self.__dict__.update(d)


class MaxEntResult(MaxEntResultData):
    """Collects the sweep records of a run (one per matrix element) and derives the result arrays."""

    def __init__(self, matrix_structure=None, element_wise=True, complex_elements=False, use_hermiticity=True):
        super(MaxEntResult, self).__init__(matrix_structure, element_wise, complex_elements, use_hermiticity)
        self._records = {}                     # element key (None or index tuple) -> dict of arrays
        self._results_from_analyzers = self._empty(dict)
        self._analysis_cache = {}
        self._start, self._end = {}, {}

    # ---- bookkeeping ------------------------------------------------------------------------------
    def _shape(self):
        if self._matrix_structure is None or not self._element_wise:
            return ()
        return self._matrix_structure + ((2,) if self._complex_elements else ())

    def _empty(self, make):
        return _nested(self._shape(), make)

    def _key(self, matrix_element, complex_index):
        if matrix_element is None:
            return None
        key = tuple(matrix_element)
        if self._complex_elements and complex_index is not None:
            key = key + (complex_index,)
        return key

    def _invalidate(self):
        self._saved = dict()
        self._analysis_cache = {}

    def add_sweep(self, record, matrix_element=None, complex_index=None):
        """Store the alpha-resolved arrays of one element.  ``record`` keys: alpha, v, chi2, S, Q, H, A,
        probability (or None), omega, G, G_orig, data_variable, G_rec, n_iter, converged."""
        self._invalidate()
        self._records[self._key(matrix_element, complex_index)] = record

    def add_result(self, solution, log_probability=None, matrix_element=None, complex_index=None):
        """Append ONE alpha (python/maxent_result.py:751-791).  ``solution`` provides the attributes
        ``alpha, v, chi2, S, Q, H, A`` and the per-element constants ``omega, G, G_orig, data_variable, G_rec``."""
        self._invalidate()
        key = self._key(matrix_element, complex_index)
        rec = self._records.setdefault(key, dict(_rows=[]))
        rec.setdefault('_rows', []).append((solution, log_probability))
        rows = rec['_rows']
        get = lambda name: np.array([getattr(s, name) for s, _ in rows])
        for name in ('alpha', 'v', 'chi2', 'S', 'Q', 'H', 'A', 'G_rec'):
            rec[name] = get(name)
        rec['probability'] = np.array([np.nan if p is None else p for _, p in rows])
        for name in ('omega', 'G', 'G_orig', 'data_variable'):
            rec[name] = getattr(solution, name)

    def analyze(self, analyzers, matrix_element=None, complex_index=None):
        """Hand the alpha-resolved data of one element to every analyzer; a ValueError is stored as its text
        (python/maxent_result.py:793-822)."""
        key = self._key(matrix_element, complex_index)
        slot = self._get_element(self._results_from_analyzers, key)
        slot.clear()
        for analyzer in analyzers:
            try:
                res = analyzer.analyze(self, key)
                res.maxent_result = self
                slot[res['name']] = res
            except ValueError as e:
                slot[analyzer.name] = str(e)
        self._saved.pop('analyzer_results', None)

    def start_timing(self, matrix_element=None, complex_index=None, time=None):
        self._start[self._key(matrix_element, complex_index)] = datetime.now() if time is None else time

    def end_timing(self, matrix_element=None, complex_index=None, time=None):
        key = self._key(matrix_element, complex_index)
        self._end[key] = datetime.now() if time is None else time
        return self._end[key] - self._start.get(key, self._end[key])

    # ---- assembling arrays ------------------------------------------------------------------------
    def _longest(self):
        """The record with the most alpha values (defines ``alpha`` and ``omega``)."""
        best = None
        for rec in self._records.values():
            if 'alpha' in rec and (best is None or len(rec['alpha']) > len(best['alpha'])):
                best = rec
        if best is None:
            raise AttributeError('no results stored yet')
        return best

    def _assemble(self, name, hermiticity_conjugate=False, dtype=float):
        shape = self._shape()
        if not shape:
            return np.array(self._records[None][name]) if None in self._records else np.array([])
        have = {k: np.asarray(r[name]) for k, r in self._records.items() if name in r and r[name] is not None}
        if not have:
            return np.full(shape + (0,), np.nan)
        ndim = max(a.ndim for a in have.values())
        trailing = tuple(max(a.shape[d] for a in have.values() if a.ndim == ndim) for d in range(ndim))
        out = np.full(shape + trailing, np.nan, dtype=dtype)
        for k, a in have.items():
            out[k + tuple(slice(0, n) for n in a.shape)] = a
        if self._use_hermiticity and hermiticity_conjugate:
            for elem in product(*map(range, self._matrix_structure)):
                if elem == elem[::-1] or not np.all(np.isnan(out[elem])):
                    continue
                out[elem] = out[elem[::-1]]
                if self._complex_elements:
                    out[elem + (1,)] = -out[elem + (1,)]
        return out

    @saved
    def alpha(self):
        """The (scaled) alpha values actually used."""
        return np.array(self._longest()['alpha'])

    @saved
    def omega(self):
        return self._longest()['omega']

    @saved
    def v(self):
        """Singular-space solution vectors, [..., n_alpha, n_sv] (NaN-padded).  Basis-dependent: the columns of
        V are those of this engine's SVD (see test/python/maxent_result.py:26-35 in the reference)."""
        return self._assemble('v')

    @saved
    def chi2(self):
        return self._assemble('chi2')

    @saved
    def S(self):
        return self._assemble('S')

    @saved
    def Q(self):
        return self._assemble('Q')

    @saved
    def H(self):
        return self._assemble('H', hermiticity_conjugate=True)

    @saved
    def A(self):
        return self._assemble('A', hermiticity_conjugate=True)

    @saved
    def probability(self):
        return self._assemble('probability')

    @saved
    def G(self):
        """Input data (rotated when a covariance was set)."""
        return self._assemble('G')

    @saved
    def G_orig(self):
        return self._assemble('G_orig')

    @saved
    def data_variable(self):
        return self._assemble('data_variable')

    @saved
    def G_rec(self):
        """Reconstructed data K_delta A for every alpha, in the original basis (python/maxent_result.py:905-908)."""
        return self._assemble('G_rec')

    @property
    def n_iter(self):
        """Levenberg iterations per alpha (device counter; not a field of the reference, not saved)."""
        return self._assemble('n_iter')

    @property
    def converged(self):
        """Per alpha: did the minimiser meet its convergence criterion (the '!' flag of the log lines)."""
        c = self._assemble('converged')
        return c.astype(bool) if not self._shape() else c

    def _per_element(self, values, default):
        shape = self._shape()
        if not shape:
            return values.get(None, default)
        out = np.empty(shape, dtype=object)
        for elem in product(*map(range, shape)):
            out[elem] = values.get(elem, default)
        return out

    @saved
    def run_time_total(self):
        d = {k: self._end[k] - self._start[k] for k in self._end if k in self._start}
        return self._per_element(d, timedelta(0))

    @saved
    def run_times(self):
        """Per-alpha run times.  The fused sweep is one launch, so its time is spread evenly over the alphas."""
        d = {}
        for k, rec in self._records.items():
            n = len(rec.get('alpha', ()))
            tot = (self._end[k] - self._start[k]) if (k in self._end and k in self._start) else timedelta(0)
            d[k] = [tot / n] * n if n else []
        return self._per_element(d, [])

    @saved
    def analyzer_results(self):
        return self._results_from_analyzers

    @saved
    def matrix_structure(self):
        return self._matrix_structure

    @saved
    def effective_matrix_structure(self):
        if self._matrix_structure is not None and self._element_wise and self._complex_elements:
            return self._matrix_structure + (2,)
        return self._matrix_structure

    @saved
    def default_analyzer_name(self):
        return self._default_analyzer_name

    @saved
    def element_wise(self):
        return self._element_wise

    @saved
    def zero_elements(self):
        return self._zero_elements

    @saved
    def use_hermiticity(self):
        return self._use_hermiticity

    @saved
    def complex_elements(self):
        return self._complex_elements

    @property
    def data(self):
        """Array-only twin (``MaxEntResultData``) that can be written to h5 or pickled."""
        try:
            return MaxEntResultData.__factory_from_dict__("MaxEntResultData", copy.copy(self.__reduce_to_dict__()))
        except AttributeError as e:
            raise Exception(e)
