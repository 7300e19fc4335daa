"""
Progress print-outs of the mirrored reference classes, silenceable per THREAD.

The reference prints its progress with plain print(). Callers that run several fits or file
preparations on different threads (fit_stream workers, run_many's prefetch thread) must be able to
silence one thread without touching sys.stdout, which contextlib.redirect_stdout swaps for the
whole process (interleaved enter / exit from two threads can leave it pointing at a dead StringIO).
"""
import contextlib
import threading

_thread_state = threading.local()


@contextlib.contextmanager
def quiet(enabled: bool = True):
    """Silences say() for the calling thread only (nestable)."""
    previous = getattr(_thread_state, "quiet", False)
    _thread_state.quiet = previous or bool(enabled)
    try:
        yield
    finally:
        _thread_state.quiet = previous


def say(*args, **kwargs):
    if not getattr(_thread_state, "quiet", False):
        print(*args, **kwargs)
