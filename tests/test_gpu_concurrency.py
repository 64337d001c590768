"""Races: the same input must give the same bytes every time, alone, on concurrent streams and from concurrent host
threads.  (compute-sanitizer's racecheck is not available on the GPU pool; a data race in an mbarrier pipeline, a
workspace shared by two streams or a process-global launch flag shows up here as a run that differs.)"""
import threading

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(qparams):
    from convnet_quantization_b200.engine import StaticEngine
    return StaticEngine(qparams, "cuda")


def _images(b, seed):
    from convnet_quantization_b200 import synth
    return synth.images_f32(b, seed=seed).cuda().contiguous()


@pytest.mark.parametrize("b,reps", [(1, 200), (33, 100), (1500, 40), (16384, 12)])
def test_repeated_forwards_are_bit_identical(engine, b, reps):
    x = _images(b, 7)
    first = engine.forward(x).clone()
    for i in range(reps):
        assert torch.equal(engine.forward(x), first), f"batch {b}: forward {i + 1} differs from the first"


def test_taps_are_bit_identical_across_runs(engine):
    """Every intermediate activation, not just the logits (a race could hide behind the arg-max-like final layers)."""
    x = _images(600, 3)
    _, first = engine.forward(x, taps=True)
    first = {k: v.clone() for k, v in first.items()}
    for _ in range(10):
        _, taps = engine.forward(x, taps=True)
        for k, v in taps.items():
            assert torch.equal(v, first[k]), k


def test_two_streams_one_engine(engine):
    """Forwards of one engine in flight on two CUDA streams at once (what the host-input pipeline does): each stream has
    its own workspace, kernels of both are co-resident on the SMs (TMEM allocation, shared memory carve-out)."""
    xs = [_images(2048, 11), _images(1536, 12)]
    want = [engine.forward(x).clone() for x in xs]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [[], []]
    torch.cuda.synchronize()
    for _ in range(12):
        for k in (0, 1):
            with torch.cuda.stream(streams[k]):
                outs[k].append(engine.forward(xs[k], graph=False))
    torch.cuda.synchronize()
    for k in (0, 1):
        for i, y in enumerate(outs[k]):
            assert torch.equal(y, want[k]), f"stream {k} forward {i}"


def test_two_host_threads(qparams):
    """Two host threads, each with its own engine and stream: one replays the small-batch graph (programmatic dependent
    launch ON for its launches), the other runs the eager large-batch path (PDL OFF) - the launch attribute is
    thread-local state in the library, the launch counter is shared."""
    from convnet_quantization_b200 import _lib
    from convnet_quantization_b200.engine import StaticEngine
    lib = _lib.load()
    cfg = [(8, True, 300), (3000, False, 40)]
    engines = [StaticEngine(qparams, "cuda", use_graphs=g) for _, g, _ in cfg]
    xs = [_images(b, 20 + i) for i, (b, _, _) in enumerate(cfg)]
    want = [e.forward(x, graph=False).clone() for e, x in zip(engines, xs)]
    torch.cuda.synchronize()
    errors = []
    start = threading.Barrier(2)

    def work(i):
        try:
            torch.cuda.set_device(0)
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                start.wait()
                for r in range(cfg[i][2]):
                    y = engines[i].forward(xs[i])
                    if r % 10 == 0 and not torch.equal(y, want[i]):
                        errors.append(f"thread {i} forward {r} differs")
                y = engines[i].forward(xs[i])
                s.synchronize()
                if not torch.equal(y, want[i]):
                    errors.append(f"thread {i} final forward differs")
        except Exception as e:  # noqa: BLE001 - reported below
            errors.append(f"thread {i}: {type(e).__name__}: {e}")

    n0 = lib.b200q_launch_count()
    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in threads), "a worker thread hung"
    assert not errors, errors
    assert lib.b200q_launch_count() > n0


def test_dynamic_linear_on_two_streams():
    """b200q_linear_dynamic keeps its reduction scratch per weights object; two objects on two streams do not interact."""
    from convnet_quantization_b200 import ops
    gen = torch.Generator().manual_seed(5)
    Ws = [ops.DynamicLinearWeights(torch.randint(-127, 128, (512, 4096), dtype=torch.int8, generator=gen), 0.01 * (i + 1),
                                   torch.randn(512, generator=gen), "cuda") for i in range(2)]
    xs = [(torch.randn(700 + 300 * i, 4096, generator=gen) * (i + 1)).cuda() for i in range(2)]
    want = [ops.linear_dynamic(x, W, relu=True).clone() for x, W in zip(xs, Ws)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [[], []]
    torch.cuda.synchronize()
    for _ in range(15):
        for k in (0, 1):
            with torch.cuda.stream(streams[k]):
                outs[k].append(ops.linear_dynamic(xs[k], Ws[k], relu=True))
    torch.cuda.synchronize()
    for k in (0, 1):
        assert all(torch.equal(y, want[k]) for y in outs[k]), f"stream {k}"
