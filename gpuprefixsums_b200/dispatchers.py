"""Python mirror of the reference's dispatcher classes, for the parity tests.

Same class names, constructor argument and public methods as
``GPUPrefixSumsCUDA/*Dispatcher.cuh`` (and as the C++ mirror in ``host/*.hpp``): the
dispatcher owns its buffer and scans it in place on all-ones input.  Methods return
the pass count / throughput instead of only printing, so pytest can assert on them.
"""
from __future__ import annotations

import torch

from . import _capi as capi
from .scan import PrefixSum, cub_scan, init_ones, validate_ones

PARTITION_SIZE = 4096  # the reference's tile (CSDL_WARPS*32*UINT4_PER_THREAD*4): test range only


class _DispatcherBase:
    algo = "csdl"
    name = ""

    def __init__(self, maxSize: int):
        self.k_maxSize = (maxSize + 3) // 4 * 4  # align16(), ...Dispatcher.cuh:168
        self.m_scan = torch.empty(self.k_maxSize, dtype=torch.int32, device="cuda")
        self.ps = PrefixSum(self.k_maxSize)

    def _dispatch(self, size: int, exclusive: bool) -> None:
        view = self.m_scan[:size]
        self.ps.scan(view, view, algo=self.algo, exclusive=exclusive)

    def _test_all(self, exclusive: bool, sizes=None, large=(26, 27, 28)) -> tuple[int, int]:
        """Reference protocol (...Dispatcher.cuh:38-114): every size in [4096, 8192] then
        2^26..2^28, all-ones input, closed-form validation.  Returns (passed, expected)."""
        sizes = range(PARTITION_SIZE, 2 * PARTITION_SIZE + 1) if sizes is None else sizes
        todo = list(sizes) + [1 << i for i in large if (1 << i) <= self.k_maxSize]
        passed = 0
        for size in todo:
            init_ones(self.m_scan[:size])
            self._dispatch(size, exclusive)
            self.ps.check()
            if validate_ones(self.m_scan[:size], exclusive) == 0:
                passed += 1
        return passed, len(todo)

    def TestAllInclusive(self, **kw):
        return self._test_all(False, **kw)

    def TestAllExclusive(self, **kw):
        return self._test_all(True, **kw)

    def BatchTimingInclusive(self, size: int, batchCount: int) -> float:
        """keys/sec; batchCount+1 iterations, first discarded, re-init before each."""
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total_ms = 0.0
        for i in range(batchCount + 1):
            init_ones(self.m_scan[:size])
            torch.cuda.synchronize()
            start.record()
            self._dispatch(size, False)
            stop.record()
            stop.synchronize()
            if i:
                total_ms += start.elapsed_time(stop)
        self.ps.check()
        return size / (total_ms / 1e3) * batchCount


class ChainedScanDecoupledLookbackDispatcher(_DispatcherBase):
    algo = "csdl"
    name = "ChainedScanDecoupledLookback"

    def __init__(self, maxSize: int, algo: str = "csdl"):
        super().__init__(maxSize)
        self.algo = algo


class ReduceThenScanDispatcher(_DispatcherBase):
    algo = "rts"
    name = "ReduceThenScan"


class EmulatedDeadlockingDispatcher(_DispatcherBase):
    algo = "csdl_fallback"
    name = "EmulatedDeadlocking"

    def __init__(self, maxSize: int):
        super().__init__(maxSize)
        self.partMask, self.maxSpinCount = 7, 4  # EmulatedDeadlocking.cuh:15-16
        self.stats = torch.zeros(4, dtype=torch.int32, device="cuda")

    def _dispatch(self, size: int, exclusive: bool) -> None:
        view = self.m_scan[:size]
        self.ps.scan_emulated_deadlock(view, view, exclusive=exclusive, stall_mask=self.partMask,
                                       max_spin=self.maxSpinCount, stats=self.stats)

    def TestAllExclusive(self, **kw):  # the reference class has no exclusive test
        raise AttributeError("EmulatedDeadlockingDispatcher has no TestAllExclusive")

    def BatchTimingInclusive(self, size: int, batchCount: int, partMask: int = 7, maxSpinCount: int = 4) -> float:
        self.partMask, self.maxSpinCount = partMask, maxSpinCount
        try:
            return super().BatchTimingInclusive(size, batchCount)
        finally:
            self.partMask, self.maxSpinCount = 7, 4


class CubDispatcher:
    def __init__(self, size: int):
        self.k_maxSize = size
        self.m_scan = torch.empty(size, dtype=torch.int32, device="cuda")

    def BatchTimingCubChainedScan(self, size: int, batchCount: int) -> float:
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total_ms = 0.0
        view = self.m_scan[:size]
        for i in range(batchCount + 1):
            init_ones(view)
            torch.cuda.synchronize()
            start.record()
            cub_scan(view, view)
            stop.record()
            stop.synchronize()
            if i:
                total_ms += start.elapsed_time(stop)
        return size / (total_ms / 1e3) * batchCount


__all__ = ["ChainedScanDecoupledLookbackDispatcher", "ReduceThenScanDispatcher",
           "EmulatedDeadlockingDispatcher", "CubDispatcher", "PARTITION_SIZE", "capi"]
