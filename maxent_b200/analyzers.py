"""Alpha analyzers: pick (or average) the one spectral function out of A_alpha(omega).

Interface of python/analyzers/*.py: ``Analyzer.analyze(maxent_result, matrix_element) -> AnalyzerResult``
(a dict with ``A_out``, ``alpha_index``, ``name``, ``info`` and analyzer-specific extras).  All five
reductions are computed by ONE device kernel, ``mx_analyze`` (csrc/mx_api.cu ``analyze_kernel``):

* LineFitAnalyzer        two-piece line fit of log chi2 vs log alpha      (linefit_analyzer.py:28-87,151-183)
* Chi2CurvatureAnalyzer  max curvature of log10 chi2 vs gamma log10 alpha (chi2_curvature_analyzer.py:25-49,101-131)
* EntropyAnalyzer        min (dS/dlog alpha)^2                            (entropy_analyzer.py:72-103)
* ClassicAnalyzer        max probability                                  (classic_analyzer.py:50-82)
* BryanAnalyzer          probability-weighted average of A_alpha          (bryan_analyzer.py:106-154)

``analyze_arrays`` is the array-level entry the batched front end and the result object share; each
``Analyzer.analyze`` asks it for the arrays of one matrix element and repackages its slot."""
import sys

import numpy as np

from . import _lib

SLOT = {"linefit": _lib.AN_LINEFIT, "chi2curv": _lib.AN_CHI2CURV, "entropy": _lib.AN_ENTROPY,
        "classic": _lib.AN_CLASSIC, "bryan": _lib.AN_BRYAN}


def analyze_arrays(alpha, chi2, S, probability, A, gamma=0.2, linefit_deg=0, average_by_integration=False):
    """Run mx_analyze for one spectrum given as host arrays; returns (alpha_index[5], A_out[5, n_omega], aux)."""
    from . import engine
    p = None if probability is None or np.all(np.isnan(probability)) else np.asarray(probability, dtype=np.float64)[None]
    idx, A_out, aux = engine.analyze(np.asarray(alpha, dtype=np.float64), np.asarray(chi2, dtype=np.float64)[None],
                                     np.asarray(S, dtype=np.float64)[None], p, np.asarray(A, dtype=np.float64)[None],
                                     gamma=gamma, linefit_deg=linefit_deg, bryan_by_integration=average_by_integration,
                                     want_aux=True)
    return idx[0].cpu().numpy(), A_out[0].cpu().numpy(), aux[0].cpu().numpy()


class AnalyzerResult(dict):
    """Result of one analyzer: keys ``A_out``, ``name``, ``info``, ``alpha_index`` (if applicable), extras."""

    def __reduce_to_dict__(self):
        return self

    @classmethod
    def __factory_from_dict__(cls, name, D):
        self = cls()
        self.update(D)
        return self

    def _get_maxent_result(self, maxent_result):
        if maxent_result is None:
            try:
                maxent_result = self.maxent_result
            except AttributeError:
                print('Please supply the keyword argument maxent_result', file=sys.stderr)
                raise
        return maxent_result

    # plot data providers (same tuples the reference's @plot_function methods return: x, y, options)
    def plot_A_out(self, maxent_result=None, **kwargs):
        r = self._get_maxent_result(maxent_result)
        return (r.omega, self['A_out'], dict(label=r'$A(\omega)$ {}'.format(self['name']), x_label=r'$\omega$',
                                             y_label=r'$A(\omega)$', log_x=False, log_y=False))

    def plot_curvature(self, maxent_result=None, **kwargs):
        r = self._get_maxent_result(maxent_result)
        return (r.alpha, self['curvature'], dict(label='curvature {}'.format(self['name']), x_label=r'$\alpha$',
                                                 y_label='curvature', log_x=True, log_y=False))

    def plot_dS_dalpha(self, maxent_result=None, **kwargs):
        r = self._get_maxent_result(maxent_result)
        return (r.alpha, self['dS_dalpha'], dict(label='dS_dalpha {}'.format(self['name']), x_label=r'$\alpha$',
                                                 y_label='dS_dalpha', log_x=True, log_y=False))

    def plot_linefit(self, maxent_result=None, element=None, **kwargs):
        r = self._get_maxent_result(maxent_result)
        idx = slice(None) if element is None else element
        la = np.log(r.alpha)
        p = self['linefit_params']
        return (r.alpha, np.column_stack((r.chi2[idx], np.exp(np.polyval(p[0], la)), np.exp(np.polyval(p[1], la)))),
                dict(label='linefit {}'.format(self['name']), x_label=r'$\alpha$', y_label='linefit',
                     log_x=True, log_y=True))


class Analyzer(object):
    """Base class; ``name`` defaults to the class name."""
    _slot = None

    def __init__(self, name=None):
        self.name = self.__class__.__name__ if name is None else name

    # parameters that mx_analyze needs from whichever analyzer is asking
    def _kernel_args(self):
        return {}

    def _device_slot(self, maxent_result, matrix_element):
        def elem(what):
            return maxent_result._get_element(what, matrix_element)
        args = dict(gamma=0.2, linefit_deg=0, average_by_integration=False)
        args.update(self._kernel_args())
        # one mx_analyze launch serves every analyzer that asks with the same parameters
        cache = getattr(maxent_result, "_analysis_cache", None)
        key = (matrix_element, args["gamma"], args["linefit_deg"], args["average_by_integration"])
        if cache is None or key not in cache:
            got = analyze_arrays(maxent_result.alpha, elem(maxent_result.chi2), elem(maxent_result.S),
                                 elem(maxent_result.probability), elem(maxent_result.A), **args)
            if cache is not None:
                cache[key] = got
        else:
            got = cache[key]
        idx, A_out, aux = got
        return int(idx[self._slot]), np.array(A_out[self._slot]), aux

    def analyze(self, maxent_result, matrix_element=None):
        raise NotImplementedError("Use a subclass of Analyzer")


class LineFitAnalyzer(Analyzer):
    _slot = _lib.AN_LINEFIT

    def __init__(self, linefit_deg=0, name=None):
        self.linefit_deg = linefit_deg
        super(LineFitAnalyzer, self).__init__(name=name)

    def _kernel_args(self):
        return dict(linefit_deg=self.linefit_deg)

    def analyze(self, maxent_result, matrix_element=None):
        k, A_out, aux = self._device_slot(maxent_result, matrix_element)
        if k < 0:
            raise ValueError('linefit: no valid break point (too few alpha values or chi2 is NaN)')
        res = AnalyzerResult()
        res['alpha_index'] = k
        p2 = [aux[2], aux[3]] if self.linefit_deg == 1 else [aux[3]]
        res['linefit_params'] = [np.array([aux[0], aux[1]]), np.array(p2)]
        res['A_out'] = A_out
        res['linefit_deg'] = self.linefit_deg
        res['name'] = self.name
        res['info'] = 'Ideal alpha (linefit): {} (= index {} zero-based)'.format(maxent_result.alpha[k], k)
        return res


class Chi2CurvatureAnalyzer(Analyzer):
    _slot = _lib.AN_CHI2CURV

    def __init__(self, gamma=0.2, name=None):
        self.gamma = gamma
        super(Chi2CurvatureAnalyzer, self).__init__(name=name)

    def _kernel_args(self):
        return dict(gamma=self.gamma)

    def analyze(self, maxent_result, matrix_element=None):
        k, A_out, aux = self._device_slot(maxent_result, matrix_element)
        n = len(maxent_result.alpha)
        if k < 0:
            raise ValueError('curvature is all NaN')
        res = AnalyzerResult()
        res['curvature'] = aux[4:4 + n]
        res['alpha_index'] = k
        res['A_out'] = A_out
        res['gamma'] = self.gamma
        res['name'] = self.name
        res['info'] = 'Ideal alpha (curvature): {} (= index {} zero-based)'.format(maxent_result.alpha[k], k)
        return res


class EntropyAnalyzer(Analyzer):
    _slot = _lib.AN_ENTROPY

    def analyze(self, maxent_result, matrix_element=None):
        k, A_out, aux = self._device_slot(maxent_result, matrix_element)
        n = len(maxent_result.alpha)
        if k < 0:
            raise ValueError('dS_dalpha is all NaN')
        res = AnalyzerResult()
        res['dS_dalpha'] = aux[4 + n:4 + 2 * n]
        res['alpha_index'] = k
        res['A_out'] = A_out
        res['name'] = self.name
        res['info'] = 'Ideal alpha (entropy): {} (= index {} zero-based)'.format(maxent_result.alpha[k], k)
        return res


class ClassicAnalyzer(Analyzer):
    _slot = _lib.AN_CLASSIC

    def analyze(self, maxent_result, matrix_element=None):
        res = AnalyzerResult()
        res['name'] = self.name
        if np.all(np.isnan(maxent_result._get_element(maxent_result.probability, matrix_element))):
            res['info'] = 'Probability not calculated. Cannot use ClassicAnalyzer.'
            return res
        k, A_out, _ = self._device_slot(maxent_result, matrix_element)
        res['alpha_index'] = k
        res['A_out'] = A_out
        res['info'] = 'Ideal alpha (classic): {} (= index {} zero-based)'.format(maxent_result.alpha[k], k)
        return res


class BryanAnalyzer(Analyzer):
    _slot = _lib.AN_BRYAN

    def __init__(self, average_by_integration=False, name=None):
        self.average_by_integration = average_by_integration
        super(BryanAnalyzer, self).__init__(name=name)

    def _kernel_args(self):
        return dict(average_by_integration=self.average_by_integration)

    def analyze(self, maxent_result, matrix_element=None):
        res = AnalyzerResult()
        res['name'] = self.name
        if np.all(np.isnan(maxent_result._get_element(maxent_result.probability, matrix_element))):
            res['info'] = 'Probability not calculated. Cannot use BryanAnalyzer.'
            return res
        _, A_out, _ = self._device_slot(maxent_result, matrix_element)
        res['A_out'] = A_out
        res['info'] = 'Bryan analyzer: average of A weighted by probability calculated.'
        return res
