"""Worker-thread state machine Init -> Run -> Stop (behaviour of /root/reference/Core/InfernWrkThread.py:28-69)."""
from threading import Lock, Thread

RTPWrkTInit = 0
RTPWrkTRun = 1
RTPWrkTStop = 2


class InfernWrkThread(Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.state_lock = Lock()
        self.state = RTPWrkTInit

    def get_state(self, locked: bool = False) -> int:
        if locked:
            return self.state
        with self.state_lock:
            return self.state

    def _set_state(self, newstate: int, expected_state=None, raise_on_error: bool = True) -> int:
        with self.state_lock:
            prev = self.state
            if expected_state is not None and prev != expected_state:
                if raise_on_error:
                    raise AssertionError(f"Unexpected state: {prev}, {expected_state} expected")
                return prev
            self.state = newstate
            return prev

    def thread_started(self):
        self._set_state(RTPWrkTRun, expected_state=RTPWrkTInit)

    def stop(self):
        prev = self._set_state(RTPWrkTStop, expected_state=RTPWrkTRun, raise_on_error=True)
        if prev == RTPWrkTRun:
            self.join()
        self._set_state(RTPWrkTInit, expected_state=RTPWrkTStop)
