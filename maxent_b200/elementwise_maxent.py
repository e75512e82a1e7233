"""Matrix-valued continuation, element by element (python/elementwise_maxent.py).

``ElementwiseMaxEnt`` owns two ``TauMaxEnt`` workers -- ``maxent_diagonal`` (normal entropy) and
``maxent_offdiagonal`` (plus-minus entropy for A_ij that may change sign) -- and one shared
``MaxEntResult`` with matrix structure.  Attributes that both workers agree on can be read and set on
the object itself (python/elementwise_maxent.py:110-146).

What is different from the reference: its ``run_diagonal / run_offdiagonal`` call ``run_element`` in a
serial Python loop (:223-268).  Here the matrix elements of one pass are collected as *jobs*
(``MaxEntLoop.snapshot``) and those that share kernel, error model and default model are continued in
ONE fused launch (``MaxEntLoop.run_jobs``); ``run_element`` remains available for single elements.

``DiagonalMaxEnt`` (:549-559) and ``PoormanMaxEnt`` (:562-653; off-diagonal default model
D_ij = sqrt(A_ii A_jj) + eps from the diagonal results) are provided on the same machinery."""
import numpy as np

from .default_models import DataDefaultModel
from .logtaker import VerbosityFlags
from .maxent_result import MaxEntResult
from .tau_maxent import TauMaxEnt


class CallableMethodCheck(object):
    """Calls a method on both workers and insists that they answer alike (python/elementwise_maxent.py:33-52)."""

    def __init__(self, name, func_diagonal, func_offdiagonal):
        self.name, self.func_diagonal, self.func_offdiagonal = name, func_diagonal, func_offdiagonal

    def __call__(self, *args, **kwargs):
        a = self.func_diagonal(*args, **kwargs)
        b = self.func_offdiagonal(*args, **kwargs)
        if np.all(a == b):
            return a
        raise Exception('Method {n} not uniquely defined. Use self.maxent_diagonal.{n} or '
                        'self.maxent_offdiagonal.{n}!'.format(n=self.name))


def _same(a, b):
    try:
        return bool(np.all(a == b))
    except Exception:
        return a is b


class ElementwiseMaxEnt(object):
    maxent_diagonal = None
    maxent_offdiagonal = None

    def __init__(self, use_hermiticity=True, use_complex=False, **kwargs):
        self.maxent_diagonal = TauMaxEnt(**kwargs)
        self.maxent_offdiagonal = TauMaxEnt(cost_function='plusminus', **kwargs)
        self.set_G_element = None
        self.determine_shape = None
        self.G_mat = None
        self.maxent_result = None
        self.use_hermiticity = use_hermiticity
        self.use_complex = use_complex

    # ---- attribute shadowing of the two workers --------------------------------------------------------
    def __getattr__(self, name):
        d = getattr(object.__getattribute__(self, 'maxent_diagonal'), name)
        o = getattr(object.__getattribute__(self, 'maxent_offdiagonal'), name)
        if hasattr(d, '__call__') and hasattr(o, '__call__'):
            return CallableMethodCheck(name, d, o)
        if _same(d, o):
            return d
        raise Exception('Element {n} not uniquely defined. Use self.maxent_diagonal.{n} or '
                        'self.maxent_offdiagonal.{n}!'.format(n=name))

    def __setattr__(self, name, value):
        if hasattr(self.maxent_diagonal, name) and hasattr(self.maxent_offdiagonal, name):
            setattr(self.maxent_offdiagonal, name, value)
            setattr(self.maxent_diagonal, name, value)
        else:
            object.__setattr__(self, name, value)

    # ---- result ---------------------------------------------------------------------------------------
    def prepare_maxent_result(self, overwrite=False):
        if self.maxent_result is None or overwrite:
            self.maxent_result = MaxEntResult(matrix_structure=self.determine_shape(self.G_mat), element_wise=True,
                                              use_hermiticity=self.use_hermiticity, complex_elements=self.use_complex)

    @property
    def shape(self):
        try:
            return self.determine_shape(self.G_mat)
        except Exception as e:
            print(e)
            raise Exception('Cannot determine shape.')

    # ---- jobs -----------------------------------------------------------------------------------------
    def _job(self, element, re=True):
        """Load one matrix element into the right worker and freeze it as a sweep job (None: skipped by
        hermiticity)."""
        i, j = element
        if i == j:
            worker, part = self.maxent_diagonal, True
            worker.logtaker.message(VerbosityFlags.ElementInfo, "Calling MaxEnt for element {i} {i}".format(i=i))
        else:
            worker, part = self.maxent_offdiagonal, re
            if self.use_hermiticity and i > j:
                worker.logtaker.message(VerbosityFlags.ElementInfo,
                                        "Element {} {} not calculated, can be determined from hermiticity".format(i, j))
                return None, worker
            worker.logtaker.message(VerbosityFlags.ElementInfo, "Calling MaxEnt for element {} {} ".format(i, j))
        self.set_G_element(worker, self.G_mat, (i, j), part)
        self.put_error(worker, self.get_error((i, j)))
        return worker.maxent_loop.snapshot(matrix_element=(i, j), complex_index=0 if re else 1), worker

    def _run_elements(self, elements, before=None):
        """Continue a list of (element, re) pairs: jobs of the same worker go through ``run_jobs`` together.
        ``before(element)`` is called ahead of loading each element (PoormanMaxEnt sets the default model there)."""
        self.prepare_maxent_result(overwrite=False)
        per_worker = {}
        for element, re in elements:
            if before is not None:
                before(element)
            job, worker = self._job(element, re)
            if job is not None:
                per_worker.setdefault(id(worker), (worker, []))[1].append(job)
        for worker, jobs in per_worker.values():
            worker.logtaker.welcome_message()
            worker.maxent_loop.run_jobs(jobs, self.maxent_result)
        return self.maxent_result

    def run_element(self, element, re=True):
        """Continue one matrix element (real or imaginary part) into the shared result."""
        return self._run_elements([(tuple(element), re)])

    def run_diagonal(self):
        self.maxent_diagonal.logtaker.message(VerbosityFlags.ElementInfo, "Calculating diagonal elements.")
        self._run_elements([((i, i), True) for i in range(self.shape[0])])
        if self.use_complex:
            for i in range(self.shape[0]):              # a Hermitian matrix has a real diagonal
                if (i, i, 1) not in self.maxent_result.zero_elements:
                    self.maxent_result.zero_elements.append((i, i, 1))
        return self.maxent_result

    def _offdiagonal_elements(self):
        parts = [True, False] if self.use_complex else [True]
        return [((i, j), re) for i in range(self.shape[0]) for j in range(self.shape[1]) if i != j for re in parts]

    def run_offdiagonal(self):
        self.maxent_offdiagonal.logtaker.message(VerbosityFlags.ElementInfo, "Calculating off-diagonal elements.")
        return self._run_elements(self._offdiagonal_elements())

    def run(self):
        """Diagonal elements first, then the off-diagonal ones."""
        self.run_diagonal()
        self.run_offdiagonal()
        return self.maxent_result

    # ---- input ----------------------------------------------------------------------------------------
    def set_G(self, G_mat, set_G_element, determine_shape):
        """Generic entry: the matrix, a function (maxent, G_mat, elem, re) that loads one element into a worker,
        and a function returning the matrix shape."""
        self.G_mat = G_mat
        self.set_G_element = set_G_element
        self.determine_shape = determine_shape
        self.maxent_result = None

    def set_G_tau(self, G_tau, *args, **kwargs):
        raise NotImplementedError("set_G_tau needs TRIQS Green-function objects; use set_G_tau_data")

    def set_G_iw(self, G_iw, *args, **kwargs):
        raise NotImplementedError("set_G_iw needs TRIQS Green-function objects; use set_G_tau_data")

    def set_G_tau_data(self, tau, G_tau, *args, **kwargs):
        """tau[T] and G_tau[M, N, T] (complex allowed with ``use_complex``)."""
        def load(maxent, G_mat, elem, re):
            g = G_mat[1][elem]
            maxent.set_G_tau_data(G_mat[0], np.real(g) if re else np.imag(g), *args, **kwargs)
        self.set_G((tau, G_tau), load, lambda G_mat: G_mat[1].shape[:2])

    def set_G_tau_filename_pattern(self, filename, dimension, tau_col=0, G_col_re=1, G_col_im=2, *args, **kwargs):
        """One text file per element; ``filename`` contains the placeholders {i} and {j}."""
        def load(maxent, G_mat, elem, re):
            maxent.set_G_tau_file(G_mat.format(i=elem[0], j=elem[1]), tau_col, G_col_re if re else G_col_im,
                                  *args, **kwargs)
        self.set_G(filename, load, lambda G_mat: dimension)

    def set_G_tau_filenames(self, filenames, tau_col=0, G_col_re=1, G_col_im=2, *args, **kwargs):
        """One text file per element given as a 2-D array of names."""
        def load(maxent, G_mat, elem, re):
            maxent.set_G_tau_file(G_mat[elem[0]][elem[1]], tau_col, G_col_re if re else G_col_im, *args, **kwargs)
        self.set_G(np.asarray(filenames), load, lambda G_mat: G_mat.shape)

    def set_error(self, error):
        """Scalar, [T] vector (same for all elements) or [M, N, T] array."""
        self.error = error
        self.error_dimension = 1
        self.put_error = lambda maxent, error: maxent.set_error(error)

    def set_cov(self, cov):
        """[T, T] covariance (same for all elements) or [M, N, T, T]."""
        self.error_dimension = 2
        self.error = cov
        self.put_error = lambda maxent, error: maxent.set_cov(error)

    def get_error(self, elem):
        if isinstance(self.error, float) or np.ndim(self.error) == 0:
            return self.error
        if np.ndim(self.error) == self.error_dimension:
            return self.error
        return self.error[elem]

    def get_tau(self):
        d = self.maxent_diagonal.get_data_variable()
        o = self.maxent_offdiagonal.get_data_variable()
        if _same(d, o):
            return d
        raise Exception('Element tau not uniquely defined. Use self.maxent_diagonal.tau or '
                        'self.maxent_offdiagonal.tau!')

    def set_tau(self, tau, update_K=True, update_chi2=True, update_Q=True, update_H_of_v=True):
        for worker in (self.maxent_diagonal, self.maxent_offdiagonal):
            worker.set_tau(tau, update_K=update_K, update_chi2=update_chi2, update_Q=update_Q,
                           update_H_of_v=update_H_of_v)

    tau = property(get_tau, set_tau)


class DiagonalMaxEnt(ElementwiseMaxEnt):
    """Only the diagonal elements."""

    def run(self):
        self.run_diagonal()
        return self.maxent_result

    def run_offdiagonal(self):
        raise TypeError('DiagonalMaxEnt cannot run for off-diagonals.')


class PoormanMaxEnt(ElementwiseMaxEnt):
    """Poor man's matrix method: off-diagonal elements use the default model
    D_ij = sqrt(A_ii A_jj) + D_add_constant built from the diagonal results of ``analyzer_offdiag_D``.
    Two phases with a barrier in between; every off-diagonal element has its own default model, passed to the
    device as one model per spectrum, so the off-diagonal pass is a single launch."""

    def __init__(self, analyzer_offdiag_D='LineFitAnalyzer', D_add_constant=1.e-6, *args, **kwargs):
        super(PoormanMaxEnt, self).__init__(*args, **kwargs)
        self.analyzer_offdiag_D = analyzer_offdiag_D
        self.D_add_constant = D_add_constant

    def run_offdiagonal(self):
        self.prepare_maxent_result(overwrite=False)
        self.maxent_offdiagonal.logtaker.message(
            VerbosityFlags.ElementInfo, "Calculating off-diagonal elements using default model from diagonal solution")
        ar = self.maxent_result.analyzer_results

        def diag_A(i):
            node = ar[i][i][0] if self.use_complex else ar[i][i]
            return node[self.analyzer_offdiag_D]['A_out']

        def set_model(element):
            i, j = element
            self.maxent_offdiagonal.set_D(DataDefaultModel(np.sqrt(diag_A(i) * diag_A(j)) + self.D_add_constant,
                                                           self.omega))
        # every element carries its own default model; they still share kernel and error model, so the whole
        # off-diagonal pass is ONE launch with per-spectrum models (MxProblem.per_spectrum_model)
        return self._run_elements(self._offdiagonal_elements(), before=set_model)
