"""Input pipeline (deeplio_b200/pipeline.py): batches arrive intact and in order through the double-buffered
host->device prefetcher; lagged scalar reads return every pushed value once, one step late."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_prefetcher_order_and_contents():
    from deeplio_b200.pipeline import DevicePrefetcher
    g = torch.Generator().manual_seed(3)
    batches = [{"a": torch.randn(4, 1 << 18, generator=g).pin_memory(),
                "b": torch.randint(0, 100, (7,), generator=g).pin_memory()} for _ in range(5)]
    seen = []
    for d in DevicePrefetcher(batches, "cuda:0"):
        # a long-running consumer of the slot: the next copy into it must wait for this step
        acc = d["a"].clone()
        for _ in range(20):
            acc = acc * 1.0000001
        seen.append((d["a"].clone(), d["b"].clone(), acc))
    torch.cuda.synchronize()
    assert len(seen) == len(batches)
    for (a, b, acc), h in zip(seen, batches):
        assert torch.equal(a.cpu(), h["a"]) and torch.equal(b.cpu(), h["b"])
        assert torch.allclose(acc.cpu(), h["a"], rtol=1e-4)


def test_prefetcher_rejects_cpu():
    from deeplio_b200.pipeline import DevicePrefetcher
    with pytest.raises(RuntimeError):
        DevicePrefetcher([], "cpu")


def test_lagged_scalar():
    from deeplio_b200.pipeline import LaggedScalar
    lag = LaggedScalar()
    out = [lag.push(torch.tensor(float(i), device="cuda:0") * 2) for i in range(5)]
    assert out == [None, 0.0, 2.0, 4.0, 6.0]
    assert lag.flush() == 8.0
