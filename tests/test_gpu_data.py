"""Device-side batch construction (numpy_transformer_b200/data.py) against the reference's host
procedure: transformer_of_image.py:97-139 (MNIST /255, shuffle, slice) and getinputs, gpt/gpt_train_potry3000.py:106-133.
Batches must be BIT-identical to what the host path uploads (integer / index work; fp32 pixels via the same rounding)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nt():
    import numpy_transformer_b200 as nt
    nt.init(0)
    nt.install()
    nt.set_gemm_mode(1)
    return nt


def test_image_batches_bit_identical_to_host_pipeline(nt):
    from numpy_transformer_b200 import DeviceArray
    from numpy_transformer_b200.data import ImageBatches
    rs = np.random.RandomState(0)
    n, B = 1000, 100
    raw = rs.randint(0, 256, (n, 28, 28)).astype(np.uint8)
    raw[0] = np.arange(784).reshape(28, 28) % 256               # every pixel value occurs
    lab = rs.randint(0, 10, n)
    loader = ImageBatches(raw, lab, B)
    assert len(loader) == 10
    datas, labels = raw / 255, lab.copy()                       # transformer_of_image.py:99-100
    img_out, lab_out = DeviceArray((B, 1, 28, 28)), DeviceArray((B,), np.int32)
    for epoch in range(2):
        np.random.seed(100 + epoch)
        k = np.arange(n)                                        # :128-133
        np.random.shuffle(k)
        datas, labels = datas[k], labels[k]
        np.random.seed(100 + epoch)
        loader.shuffle()                                        # same draw from the same generator state
        for j in (0, 3, 9):
            loader.batch_into(j, img_out, lab_out)
            host = datas[j * B:(j + 1) * B][:, np.newaxis].astype(np.float32)   # what the H2D cast would upload
            assert np.array_equal(img_out.numpy(), host)
            assert np.array_equal(lab_out.numpy(), labels[j * B:(j + 1) * B])
    with pytest.raises(AssertionError):
        loader.batch_into(10, img_out, lab_out)


def test_token_windows_match_getinputs(nt):
    from numpy_transformer_b200 import DeviceArray
    from numpy_transformer_b200.data import TokenWindows
    rs = np.random.RandomState(1)
    T, B, V = 16, 8, 50
    ids = rs.randint(0, V, 500)
    inputid = np.arange(0, len(ids) - T - 1)
    tw = TokenWindows(ids, T, B)
    x, y = DeviceArray((B, T), np.int32), DeviceArray((B, T), np.int32)
    np.random.seed(9)
    starts = tw.sample_into(inputid, x, y)
    np.random.seed(9)
    ref = np.random.choice(inputid, B, replace=False)           # gpt_train_potry3000.py:111
    assert np.array_equal(starts, ref)
    want = np.array([ids[s:s + T + 1] for s in ref])
    assert np.array_equal(x.numpy(), want[:, :-1]) and np.array_equal(y.numpy(), want[:, 1:])


def test_trainstep_fed_from_device_matches_host_batches(nt):
    """Same trajectory whether a step's batch comes through TrainStep.step(host arrays) or from the resident data set."""
    from numpy_transformer_b200 import trainer
    from numpy_transformer_b200.data import ImageBatches
    rs = np.random.RandomState(3)
    n, B = 64, 16
    raw, lab = rs.randint(0, 256, (n, 28, 28)).astype(np.uint8), rs.randint(0, 10, n)
    losses = []
    for mode in ("host", "device"):
        layers = trainer.build_vit_layers(B, embed_dim=32, heads=2, blocks=2, seed=0)
        ts = trainer.TrainStep(layers, "vit", (B, 1, 28, 28), lr=0.05)
        loader = ImageBatches(raw, lab, B)
        run = []
        for j in range(len(loader)):
            if mode == "host":
                run.append(ts.step((raw[j * B:(j + 1) * B] / 255)[:, np.newaxis], lab[j * B:(j + 1) * B]))
            else:
                loader.feed(j, ts)
                ts.run_device()
                run.append(float(ts.loss))
        losses.append(run)
    assert np.allclose(losses[0], losses[1], rtol=1e-6, atol=0)
