"""Queue -> batch worker thread (behaviour of /root/reference/Cluster/InfernBatchedWorker.py:7-45): block for the
first item, then drain without waiting up to max_batch_size; None stops the thread."""
from __future__ import annotations

from abc import ABC, abstractmethod
from queue import Empty, Queue
from typing import List, Optional

from infernos_b200.Core.InfernWrkThread import InfernWrkThread, RTPWrkTRun


class InfernBatchedWorker(InfernWrkThread, ABC):
    max_batch_size: int

    def __init__(self):
        super().__init__()
        self.inf_queue: "Queue[Optional[object]]" = Queue()

    def infer(self, wi: object):
        self.inf_queue.put(wi)

    def next_batch(self, block: bool = True, limit: Optional[int] = None) -> Optional[List[object]]:
        """block=False never waits (continuous batching polls while work is in flight); limit caps the batch below max_batch_size."""
        batch: List[object] = []
        cap = self.max_batch_size if limit is None else min(limit, self.max_batch_size)
        while len(batch) < cap:
            try:
                wi = self.inf_queue.get() if (block and not batch) else self.inf_queue.get_nowait()
            except Empty:
                break
            if wi is None:
                return None
            batch.append(wi)
        return batch

    @abstractmethod
    def process_batch(self, wis: List[object]):
        ...

    def run(self):
        super().thread_started()
        while self.get_state() == RTPWrkTRun:
            wis = self.next_batch()
            if wis is None:
                break
            for wi in wis:
                cb = getattr(wi, "_proc_start_cb", None)
                if cb is not None:
                    cb()
            self.process_batch(wis)

    def stop(self):
        self.inf_queue.put(None)
        super().stop()
