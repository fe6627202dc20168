"""Host placement helper of the ingest path: cpulist parsing and that the affinity block always restores the process mask."""
import os

from harp_b200 import hostmem


def test_cpulist_parsing():
    assert hostmem._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostmem._parse_cpulist("5") == {5}
    assert hostmem._parse_cpulist("") == set()


def test_affinity_block_restores_mask(monkeypatch):
    before = os.sched_getaffinity(0)
    some = {min(before)}
    monkeypatch.setattr(hostmem, "gpu_numa_node", lambda device: (0, some))
    with hostmem.gpu_numa_affinity("cuda:0") as info:
        assert info == {"numa_node": 0, "cpus": 1}
        assert os.sched_getaffinity(0) == some
    assert os.sched_getaffinity(0) == before
    monkeypatch.setattr(hostmem, "gpu_numa_node", lambda device: (None, None))  # unknown topology: nothing changes
    with hostmem.gpu_numa_affinity("cuda:0") as info:
        assert info["numa_node"] is None and os.sched_getaffinity(0) == before
