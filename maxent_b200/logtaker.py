"""Run log: verbosity bit mask + a sink that writes to stdout and (optionally) a log file.

Public surface and message semantics of python/logtaker.py (VerbosityFlags :25-79, Logtaker :82-230):
a message is emitted when ALL bits of its verbosity are enabled; "one-line" categories (solver
details) overwrite the current terminal line; errors are remembered.  On the fused path the
alpha-loop lines are printed post hoc from the device counters (same format as
python/maxent_loop.py:248-255)."""
from __future__ import print_function

import datetime
import warnings


class VerbosityFlags(object):
    """Bit mask selecting which categories of messages are shown (combine with ``|``)."""
    Quiet = 0
    Header = 1
    ElementInfo = 2
    Timing = 4
    AlphaLoop = 8
    SolverDetails = 16
    Errors = 32
    Default = Header | ElementInfo | Timing | AlphaLoop | Errors


def _enabled(mask, wanted):
    return (mask & wanted) == wanted


class Logtaker(object):

    def __init__(self):
        self.verbose = VerbosityFlags.Default
        self.one_line = VerbosityFlags.SolverDetails       # categories that rewrite the current line
        self.logfile = None
        self.logfile_verbose = None                        # None: same mask as the terminal
        self.end = '\n'
        self._error_log = []
        self._welcome_message_printed = False

    # ---- sinks -----------------------------------------------------------------------------------
    def open_logfile(self, name, append=True):
        self.logfile = open(name, 'a' if append else 'w')

    def close_logfile(self):
        if self.logfile is not None:
            self.logfile.close()
        self.logfile = None

    def message(self, message_verbosity, msg, *args, **kwargs):
        text = msg.format(*args, **kwargs)
        if _enabled(self.verbose, message_verbosity):
            if message_verbosity & self.one_line:
                self.end = ''
                print('\r', end='')
            elif self.end != '\n':
                self.end = '\n'
                print('')
            print(text, end=self.end)
        if self.logfile is not None:
            mask = self.verbose if self.logfile_verbose is None else self.logfile_verbose
            if _enabled(mask, message_verbosity):
                self.logfile.write(text + '\n')

    def error_message(self, msg, *args, **kwargs):
        self._error_log.append(msg.format(*args, **kwargs))
        self.message(VerbosityFlags.Errors, 'ERROR: ' + msg, *args, **kwargs)

    def get_error_messages(self):
        return self._error_log

    def clear_error_messages(self):
        del self._error_log[:]

    def log_time(self, message_verbosity=VerbosityFlags.Header):
        self.message(message_verbosity, str(datetime.datetime.now()))

    def welcome_message(self, always=False, message_verbosity=VerbosityFlags.Header):
        if self._welcome_message_printed and not always:
            return
        self.log_time(message_verbosity=message_verbosity)
        self.message(message_verbosity, "MaxEnt run")
        self.message(message_verbosity, "maxent_b200: B200-native engine with the TRIQS/maxent interface")
        self.message(message_verbosity,
                     "Please cite TRIQS/maxent and the appropriate original papers (see its documentation).\n")
        self._welcome_message_printed = True

    def verbosity_message(self, msg, iserr=False, *args, **kwargs):
        raise NotImplementedError('The function verbosity_message was removed. Please use message instead.')

    def logged_message(self, msg, *args, **kwargs):
        warnings.warn("logged_message is deprecated. Use message instead.", DeprecationWarning)
        self.message(0, msg, *args, **kwargs)

    def solver_verbose_callback(self, *args, **kwargs):
        self.message(VerbosityFlags.SolverDetails, *args, **kwargs)
