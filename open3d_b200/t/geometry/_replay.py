"""Transparent capacity recovery for the fused, host-sync-free integrate path.

The reference grows the block hash map inside HashMap::Activate (core/hashmap/HashMap.cpp:166-181),
synchronously, so ``VoxelBlockGrid.integrate`` never fails for lack of capacity.  The fused path
(``o3db_vbg_integrate_frame``, include/open3d_b200.h) runs up to two frames ahead of the host and
sizes the map ahead of need; a frame that still does not fit is dropped whole on the device together
with every later one, and a later call returns ``O3DB_ERR_CAPACITY`` with ``fused frame #K needed N
blocks``.  This module turns that contract back into the reference's behaviour: it keeps the last
few submitted frames (by reference — the images must not be overwritten in place until two more
frames went in, which holds for frames read from disk or a camera as in
examples/python/t_reconstruction_system/dense_slam.py), and on that error reserves, resubmits from
frame K and carries on.  No device work happens here; the class only orders calls, so that it can be
tested without a GPU.
"""
from __future__ import annotations

import re
from collections import deque

_NEEDED = re.compile(r"fused frame #(\d+) needed (\d+) blocks")

# how far the host may run ahead of the device's status (o3db_vbg_integrate_frame reads frame F-2's status
# when it submits frame F) plus one spare
RING_DEPTH = 4
# a resubmitted frame can outgrow the map again (the frames after the synchronously sized one are asynchronous
# again); every recovery at least doubles the capacity, so a handful is plenty before giving up
MAX_RECOVERIES = 8


def parse_dropped_frame(message: str):
    """(K, N) of "fused frame #K needed N blocks", or None."""
    m = _NEEDED.search(message or "")
    return (int(m.group(1)), int(m.group(2))) if m else None


class FrameReplay:
    """submit(frame): hand one frame to ``raw_submit`` (raises on error) and remember it;
    guard(fn): run any other call that can surface a dropped frame (size, frustum blocks, ray cast).
    ``is_capacity_error(exc)`` tells the capacity error from the others; ``reserve(n)`` must also re-arm the
    handle (o3db_vbg_reserve does); ``capacity()`` is the current capacity in blocks."""

    def __init__(self, raw_submit, reserve, capacity, is_capacity_error, depth=RING_DEPTH):
        self._raw_submit = raw_submit
        self._reserve = reserve
        self._capacity = capacity
        self._is_capacity_error = is_capacity_error
        self._ring = deque(maxlen=depth)     # (library frame index, frame)
        self._next = 0                       # the library's count of fused frames of this handle
        self.enabled = True
        self.recoveries = 0                  # for tests and logs

    def _submit_one(self, frame):
        self._raw_submit(*frame)
        self._ring.append((self._next, frame))
        self._next += 1

    def submit(self, *frame):
        try:                                 # the common case costs one call and one append
            self._submit_one(frame)
        except Exception as err:  # noqa: BLE001
            self._run(("frame", frame), err)

    def guard(self, fn):
        try:
            return fn()
        except Exception as err:  # noqa: BLE001
            return self._run(("call", fn), err)

    def _run(self, last, err):
        """`last` just failed with `err`: recover if that is the documented capacity error, then finish `last`."""
        replay = self._recover(err) if self.enabled else None
        if replay is None:
            raise err
        queue = deque([("frame", f) for f in replay] + [last])
        result = None
        budget = MAX_RECOVERIES - 1
        while queue:
            kind, item = queue[0]
            try:
                if kind == "frame":
                    self._submit_one(item)
                else:
                    result = item()
            except Exception as err:  # noqa: BLE001 - re-raised unless it is the documented capacity error
                replay = self._recover(err) if self.enabled and budget > 0 else None
                if replay is None:
                    raise
                budget -= 1
                queue.extendleft(("frame", f) for f in reversed(replay))
                continue
            queue.popleft()
        return result

    def _recover(self, err):
        """Frames to resubmit (oldest first) after reserving, or None when the error is not ours to handle."""
        if not self._is_capacity_error(err):
            return None
        parsed = parse_dropped_frame(str(err))
        if parsed is None:
            return None
        first, needed = parsed
        held = [i for i, _ in self._ring]
        if first >= self._next or first not in held:
            return None                      # the dropped frame is older than the ring: the caller has to resubmit
        replay = [f for i, f in self._ring if i >= first]
        kept = [(i, f) for i, f in self._ring if i < first]
        self._reserve(max(2 * int(self._capacity()), needed + max(needed // 2, 2048)))
        self._ring.clear()
        self._ring.extend(kept)
        self.recoveries += 1
        return replay
