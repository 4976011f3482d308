"""Device residency of the agent table and the results table.

``sim.people`` (a :class:`core.LaserFrame`) keeps the reference's host-visible
numpy columns -- tests and user scripts mutate them between construction and
``run()`` and read them back afterwards by agent index.
``DeviceState`` is the HBM twin those columns are copied into at ``run()``
entry and copied back from at exit: one torch CUDA buffer per column, length
``capacity``, reference dtypes, agents in the reference's order (initial
population node-contiguous, newborn cohorts appended node-major), so a host
index and a device slot are the same number and no permutation is needed.

The ``[nt, nodes(, strains)]`` int32 result arrays live on the device for the
whole run (kernels write row ``t`` directly) and come back in one bulk copy.

PyTorch owns every allocation; liblpk only borrows raw pointers.
"""

from __future__ import annotations

import numpy as np
import torch

# results entries that are not [nt, nodes(, strains)] int32 rows
HOST_ONLY_RESULTS = ("network",)
# agent columns no kernel writes (the host copy stays current), and columns only births write (appended cohorts)
READ_ONLY_COLUMNS = frozenset({"chronically_missed", "acq_risk_multiplier", "daily_infectivity"})
APPEND_ONLY_COLUMNS = frozenset({"node_id", "date_of_birth", "date_of_death"})


def _to_dev(arr: np.ndarray, device) -> torch.Tensor:
    t = torch.from_numpy(arr)
    return t.to(device, non_blocking=t.is_pinned())


class DeviceState:
    def __init__(self, sim, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("laser_polio_b200 runs its per-tick path on a CUDA device; none is available (no CPU fallback)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.sim = sim
        self.n_nodes = len(sim.nodes)
        self.n_strains = len(sim.pars.strain_ids)
        self.cols: dict[str, torch.Tensor] = {}
        self.res: dict[str, torch.Tensor] = {}
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.dirty: set[str] = set()  # columns a custom component wrote on the device although the stock kernels never do
        self.upload()

    # ------------------------------------------------------------------ transfers
    def upload(self):
        people, results = self.sim.people, self.sim.results
        self.count0 = int(people.count)  # slots in use at upload: append-only columns come back from here on
        columns = people.columns()
        for name, col in columns.items():
            self.cols[name] = _to_dev(col, self.device)
            self.h2d_bytes += col.nbytes
        for name, arr in results.__dict__.items():
            if isinstance(arr, np.ndarray) and arr.dtype == np.int32 and arr.ndim >= 2 and name not in HOST_ONLY_RESULTS:
                self.res[name] = _to_dev(arr, self.device)
                self.h2d_bytes += arr.nbytes
        n, ns, dev = self.n_nodes, self.n_strains, self.device
        self.zero_pop = torch.zeros(n, dtype=torch.int32, device=dev)
        # live-slot counters {agents when the previous tick ended, agents now}: births are created on the device
        self.counts = torch.tensor([people.count, people.count], dtype=torch.int64, device=dev)
        self.status = torch.zeros(2, dtype=torch.int32, device=dev)  # lpk_births_args.status: flag bits, tick of the first overflow
        self.cohort_ws = torch.zeros(2, dtype=torch.int64, device=dev)
        self.node_offsets_ws = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        self.tile_node = None  # owned by engine.FusedEngine when ticks are fused
        self.scratch_i32 = [torch.zeros(n, dtype=torch.int32, device=dev) for _ in range(4)]
        from ._lpk import RISK_BINS

        self.tally = (torch.zeros((n, ns), dtype=torch.int64, device=dev), torch.zeros(n, dtype=torch.int64, device=dev),
                      torch.zeros(n, dtype=torch.int64, device=dev), torch.zeros((n, RISK_BINS), dtype=torch.int32, device=dev))
        self.node_out = (torch.zeros(n, dtype=torch.float32, device=dev), torch.zeros((n, ns), dtype=torch.float64, device=dev),
                         torch.zeros((n, ns), dtype=torch.float64, device=dev), torch.zeros(n, dtype=torch.float64, device=dev),
                         torch.zeros(2 * n, dtype=torch.float64, device=dev))
        self.n_new = torch.zeros((n, ns), dtype=torch.int32, device=dev)
        mk = lambda *s: torch.zeros(s, dtype=torch.int32, device=dev)  # noqa: E731
        self.census = (mk(n), mk(n), mk(n), mk(n), mk(n, ns), mk(n, ns), mk(n), mk(n))
        self._net_src = None
        self.network = None
        # compaction (pars.compact_every): slots [0, counts[1]) live + recently dead, [counts[1], cap_eff) unborn pre-drawn slots,
        # [cap_eff, capacity) the graveyard (dead agents moved out of the swept range); orig[slot] = the agent's index in the
        # reference's order (None: identity, never compacted)
        self.cap_eff = int(people.capacity)
        self.graveyard = 0
        self.orig = None

    def mark_dirty(self, name: str):
        """A custom component wrote column ``name`` on the device: bring it back whole at download()."""
        self.dirty.add(name)

    def download(self):
        """Bulk D2H of every agent column the device may have changed and every device-resident results array, in
        place.  Columns no kernel writes (risk, infectivity, chronically_missed) are not copied -- the host arrays
        they were uploaded from are still current -- and columns only births write (node_id, date_of_birth,
        date_of_death) come back for the cohorts appended since upload()."""
        people, results = self.sim.people, self.sim.results
        torch.cuda.current_stream().synchronize()
        count = self.sync_count(check=False)  # a cohort that did not fit raises AFTER everything has been copied back
        inv = None
        if self.orig is not None:  # compacted: slot s holds the agent the reference keeps at index orig[s]
            inv = torch.empty_like(self.orig, dtype=torch.int64)
            inv[self.orig.long()] = torch.arange(self.orig.numel(), dtype=torch.int64, device=self.device)
        for name, t in self.cols.items():
            host = getattr(people, name)
            if inv is not None:
                if name in READ_ONLY_COLUMNS and name not in self.dirty:
                    continue  # permuted on the device, unchanged in value: the host copy is current, in the reference's order
                torch.from_numpy(host).copy_(t[inv], non_blocking=False)
                self.d2h_bytes += host.nbytes
                continue
            if name in self.dirty:
                pass
            elif name in READ_ONLY_COLUMNS:
                continue
            elif name in APPEND_ONLY_COLUMNS:
                lo, hi = min(self.count0, count), count
                if hi > lo:
                    torch.from_numpy(host[lo:hi]).copy_(t[lo:hi], non_blocking=False)
                    self.d2h_bytes += host[lo:hi].nbytes
                continue
            torch.from_numpy(host).copy_(t, non_blocking=False)
            self.d2h_bytes += host.nbytes
        for name, t in self.res.items():
            host = getattr(results, name)
            torch.from_numpy(host).copy_(t, non_blocking=False)
            self.d2h_bytes += host.nbytes
        self.release_network()
        self.check_status()

    def push_rows(self, name: str, start: int, end: int):
        """H2D of a slice of one agent column (newborn cohort written on the host)."""
        host = getattr(self.sim.people, name)[start:end]
        self.cols[name][start:end].copy_(torch.from_numpy(host), non_blocking=False)
        self.h2d_bytes += host.nbytes

    def pop_row(self, t: int) -> torch.Tensor:
        """results.pop[t] on the device (all zeros when nothing maintains it, like the reference's untouched rows)."""
        pop = self.res.get("pop")
        return self.zero_pop if pop is None else pop[t]

    def check_status(self, status=None):
        """Raise what LaserFrame.add raises in the reference when a cohort does not fit (the device flags it and creates no
        further cohort: lpk_births_args.status)."""
        flags, tick = (int(v) for v in (self.status.tolist() if status is None else status))
        if flags & 1:
            raise ValueError(f"frame.add() exceeds capacity (capacity={self.sim.people.capacity}) at tick {tick}: "
                             "that cohort and every later one were not created")
        if flags & 2:
            raise ValueError("an acq_risk_multiplier exceeds the range of the agenda's risk code")

    def sync_count(self, check=True) -> int:
        """Device -> host: how many slots are in use on the device (blocks on the stream), and whether a cohort overflowed
        capacity.  people.count is the reference's count: every agent ever created, i.e. device slots in use + graveyard."""
        count = int(self.counts[1].item())
        self.sim.people._count = count + self.graveyard
        if check:
            self.check_status()
        return count

    def set_count(self, count: int):
        """``count`` in the reference's sense (people.count)."""
        count = int(count) - self.graveyard
        self.counts.copy_(torch.tensor([count, count], dtype=torch.int64))

    # ------------------------------------------------------------------ compaction (north star: free-slot reuse + compaction)
    def warm_compaction(self):
        """One throw-away sort + gather of the table's size, so that the caching allocator already holds the multi-GB scratch a
        compaction needs (the first cudaMalloc of it costs ~0.5 s at 2.2e8 agents, against ~40 ms for the compaction itself)."""
        key = torch.zeros(self.cap_eff, dtype=torch.int16, device=self.device)
        perm = torch.sort(key, stable=True).indices
        _ = self.cols["date_of_death" if "date_of_death" in self.cols else "node_id"][:self.cap_eff][perm]
        del key, perm, _

    def compact(self):
        """Stable re-sort of the canonical table: live agents by node (previous order kept within a node), then the unborn
        pre-drawn slots in their order, then the agents that died since the last compaction, which join the graveyard at the end
        of the table and are no longer swept.  Every column moves together, and ``orig`` with them, so to_host() can hand the
        columns back in the reference's order.  The reference appends forever and scans its tombstones on every tick
        (model.py:1719-1732); this is the table operation the north star asks for instead.  Returns (live, newly dead).

        A maintenance operation, not a per-tick kernel: one radix sort of int16 keys (torch.sort) and one gather per column."""
        n, cap_eff = self.n_nodes, self.cap_eff
        count = int(self.counts[1].item())
        if self.orig is None:
            self.orig = torch.arange(self.sim.people.capacity, dtype=torch.int32, device=self.device)
        state, nid = self.cols["disease_state"], self.cols["node_id"]
        key = torch.full((cap_eff,), n, dtype=torch.int16, device=self.device)  # unborn slots: after every node
        key[:count] = torch.where(state[:count] >= 0, nid[:count], torch.full_like(nid[:count], n + 1))  # the dead: last
        live = int((key[:count] < n).sum().item())
        dead = count - live
        perm = torch.sort(key, stable=True).indices
        del key
        for t in list(self.cols.values()) + [self.orig]:
            t[:cap_eff] = t[:cap_eff][perm]
        del perm
        self.cap_eff -= dead
        self.graveyard += dead
        self.counts.copy_(torch.tensor([live, live], dtype=torch.int64))
        return live, dead


    def network_tensor(self, host_network) -> torch.Tensor:
        """Device copy of ``tx.network`` (float64, row-major); re-uploaded when the host object is replaced
        (the reference re-reads the attribute every tick, model.py:1335, and tests overwrite it)."""
        if self._net_src is not host_network:
            arr = np.ascontiguousarray(np.asarray(host_network, dtype=np.float64))
            if arr.shape != (self.n_nodes, self.n_nodes):
                raise ValueError(f"network must be {self.n_nodes}x{self.n_nodes}, got {arr.shape}")
            self.network = torch.from_numpy(arr).to(self.device)
            self.release_network()
            self._net_src = host_network
            # the reference re-reads tx.network every tick, so an in-place edit takes effect there; here the device holds a
            # copy, so while it does the host array is read-only: an in-place edit raises instead of being silently ignored
            # (assign a new array to tx.network to change it; the identity check above picks that up)
            # (only arrays that can be unlocked again: numpy refuses to re-enable writing on a view of foreign memory)
            if isinstance(host_network, np.ndarray) and host_network.flags.writeable and \
                    (host_network.flags.owndata or isinstance(host_network.base, np.ndarray)):
                host_network.setflags(write=False)
                self._net_locked = True
            self.h2d_bytes += arr.nbytes
        return self.network

    def release_network(self):
        """Give the host network array back its write permission (download(), or a new array took its place)."""
        if getattr(self, "_net_locked", False) and isinstance(self._net_src, np.ndarray):
            try:
                self._net_src.setflags(write=True)
            except ValueError:  # a view whose owner is itself read-only: leave it
                pass
        self._net_locked = False
