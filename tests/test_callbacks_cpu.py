"""Host logic of the epoch callbacks (multiplanarunet_b200/callbacks): Keras-2.3 semantics of ReduceLROnPlateau /
EarlyStopping / CSVLogger as the reference's YAML uses them, ModelCheckPointClean's file handling
(mpunet/callbacks/mcp_clean.py), FGBatchBalancer, DelayedCallback, the descriptor helpers
(mpunet/callbacks/funcs.py).  No device work."""
import csv
import os

import numpy as np
import pytest

from multiplanarunet_b200 import callbacks as C


class _Opt:
    lr = 1e-3


class _Model:
    def __init__(self):
        self.optimizer = _Opt()
        self.stop_training = False
        self.saved = []

    def save_weights(self, path):
        self.saved.append(path)
        open(path, "w").write("w")


def test_reduce_lr_on_plateau_preset():
    m = _Model()
    cb = C.ReduceLROnPlateau(patience=2, factor=0.90, verbose=0, monitor="val_dice", mode="max")
    cb.set_model(m)
    cb.on_train_begin()
    seq = [0.5, 0.6, 0.6, 0.60005, 0.59, 0.7, 0.1, 0.1]
    lrs = []
    for e, v in enumerate(seq):
        logs = {"val_dice": v}
        cb.on_epoch_end(e, logs)
        lrs.append(m.optimizer.lr)
        assert "lr" in logs
    # improvements must exceed min_delta=1e-4: epochs 2,3 do not improve -> reduce at epoch 3 (wait == patience)
    assert lrs[:3] == [1e-3] * 3 and abs(lrs[3] - 9e-4) < 1e-12
    assert abs(lrs[4] - 9e-4) < 1e-12 and abs(lrs[5] - 9e-4) < 1e-12      # wait restarted, then improvement
    assert abs(lrs[7] - 8.1e-4) < 1e-12
    with pytest.raises(ValueError):
        C.ReduceLROnPlateau(factor=1.0)
    cb2 = C.ReduceLROnPlateau(monitor="val_loss", patience=1, factor=0.5, min_lr=4e-4, min_delta=0)
    cb2.set_model(m)
    m.optimizer.lr = 1e-3
    cb2.on_train_begin()
    for e, v in enumerate([1.0, 1.0, 1.0, 1.0]):
        cb2.on_epoch_end(e, {"val_loss": v})
    assert m.optimizer.lr == 4e-4     # clipped at min_lr, mode auto -> min for a loss


def test_early_stopping_preset():
    m = _Model()
    cb = C.EarlyStopping(monitor="val_dice", min_delta=0, patience=3, verbose=1, mode="max")
    cb.set_model(m)
    cb.on_train_begin()
    for e, v in enumerate([0.3, 0.4, 0.4, 0.39, 0.4]):
        cb.on_epoch_end(e, {"val_dice": v})
        assert m.stop_training == (e == 4)
    assert cb.stopped_epoch == 4 and cb.best == 0.4


def test_model_checkpoint_clean_and_csv(tmp_path):
    m = _Model()
    os.makedirs(tmp_path / "model")
    cb = C.ModelCheckPointClean(filepath=str(tmp_path / "model" / "@epoch_{epoch:02d}_val_dice_{val_dice:.5f}.h5"),
                                monitor="val_dice", save_best_only=True, save_weights_only=True, verbose=0,
                                mode="max")
    cb.set_model(m)
    csvcb = C.CSVLogger(filename=str(tmp_path / "logs" / "training.csv"), separator=",", append=True)
    csvcb.on_train_begin()
    for e, v in enumerate([0.2, 0.5, 0.4]):
        logs = {"val_dice": v, "loss": 1.0 - v}
        cb.on_epoch_end(e, logs)
        csvcb.on_epoch_end(e, logs)
    csvcb.on_train_end()
    files = sorted(os.listdir(tmp_path / "model"))
    assert files == ["@epoch_01_val_dice_0.50000.npz"]            # older best removed, 0-based epoch in the name
    rows = list(csv.DictReader(open(tmp_path / "logs" / "training.csv")))
    assert [r["epoch"] for r in rows] == ["0", "1", "2"] and set(rows[0]) == {"epoch", "loss", "val_dice"}
    # appending continues without a second header
    csv2 = C.CSVLogger(filename=str(tmp_path / "logs" / "training.csv"), append=True)
    csv2.on_train_begin()
    csv2.on_epoch_end(3, {"val_dice": 0.6, "loss": 0.4})
    csv2.on_train_end()
    assert len(open(tmp_path / "logs" / "training.csv").read().strip().splitlines()) == 5


def test_fg_balancer_delayed_and_descriptors():
    class Seq:
        fg_batch_fraction = 0.5
        batch_size = 16

        @property
        def n_fg_slices(self):
            return int(np.ceil(self.batch_size * self.fg_batch_fraction))
    tr = Seq()
    msgs = []
    fb = C.FGBatchBalancer(tr, logger=msgs.append)
    fb.on_epoch_end(0, {"val_recall": 0.8})
    assert abs(tr.fg_batch_fraction - 0.2) < 1e-12 and tr.n_fg_slices == 4
    fb.on_epoch_end(1, {"val_recall": 1.0})
    assert tr.fg_batch_fraction == 0.01
    fb.on_epoch_end(2, {})
    assert not fb.active
    hits = []

    class Probe(C.Callback):
        def on_epoch_end(self, epoch, logs=None):
            hits.append(epoch)
    d = C.DelayedCallback(Probe(), start_from=3, logger=msgs.append)
    for e in range(5):
        d.on_epoch_end(e)
    assert hits == [2, 3, 4]
    descr = [dict(class_name="ReduceLROnPlateau", kwargs=dict(patience=2, factor=0.9, monitor="val_dice", mode="max")),
             dict(class_name="TrainTimer", pass_logger=True, kwargs=dict(verbose=True)),
             dict(class_name="EarlyStopping", kwargs=dict(monitor="val_dice", patience=15, mode="max"), start_from=4),
             dict(class_name="CSVLogger", kwargs=dict(filename="logs/training.csv", separator=",", append=True))]
    kept = [dict(d_) for d_ in descr]
    C.remove_validation_callbacks(kept, msgs.append)
    assert [k["class_name"] for k in kept] == ["TrainTimer", "CSVLogger"]
    objs, by_name = C.init_callback_objects([dict(d_) for d_ in descr], msgs.append)
    assert isinstance(objs[2], C.DelayedCallback) and isinstance(by_name["TrainTimer"], C.TrainTimer)
    with pytest.raises(ValueError):
        C.init_callback_objects([dict(class_name="NoSuchCallback", kwargs={})], msgs.append)


def test_continue_training_csv_helpers(tmp_path):
    """utils.get_last_epoch / get_lr_at_epoch / clear_csv_after_epoch (mpunet/utils/utils.py:133-177)."""
    from multiplanarunet_b200.utils import utils as U
    logs = tmp_path / "logs"
    os.makedirs(logs)
    f = logs / "training.csv"
    assert U.get_last_epoch(str(f)) == 0 and U.get_lr_at_epoch(3, str(logs)) == (None, None)
    with open(f, "w") as fh:
        fh.write("epoch,loss,lr\n0,1.0,0.001\n1,0.9,0.001\n0,1.1,0.002\n1,0.8,0.002\n2,0.7,0.0018\n3,0.6,0.0018\n")
    assert U.get_last_epoch(str(f)) == 3
    U.clear_csv_after_epoch(2, str(f))                      # keeps the last run, epochs 0..2
    rows = list(csv.DictReader(open(f)))
    assert [r["epoch"] for r in rows] == ["0", "1", "2"] and rows[0]["lr"] == "0.002"
    assert U.get_lr_at_epoch(2, str(logs)) == (0.0018, "lr")
    assert U.get_lr_at_epoch(7, str(logs)) == (None, None)


def test_fit_loop_drives_callbacks_and_stops():
    """train.fit_loop (the loop Keras' model.fit runs for the reference, train/trainer.py:246-257) with a fake
    model: callback order, logs/history, early stop, stop-flag synchronisation hook."""
    from multiplanarunet_b200.train import fit_loop
    m = _Model()
    seen = []
    m.train_on_batch = lambda x, y, w: seen.append(x) or 1.0 / (len(seen))

    class Val(C.Callback):
        def on_epoch_end(self, epoch, logs=None):
            logs["val_dice"] = [0.2, 0.3, 0.3, 0.3, 0.9][epoch]
    order = []

    class Probe(C.Callback):
        def on_train_begin(self, logs=None):
            order.append("begin")

        def on_epoch_begin(self, epoch, logs=None):
            order.append("eb%d" % epoch)

        def on_epoch_end(self, epoch, logs=None):
            order.append("ee%d" % epoch)
            assert "val_dice" in logs and "loss" in logs       # Validation ran first

        def on_train_end(self, logs=None):
            order.append("end")
    es = C.EarlyStopping(monitor="val_dice", patience=2, mode="max")
    batches = ((i, None, None) for i in range(1000))
    msgs = []
    hist = fit_loop(m, batches, steps_per_epoch=3, epochs=10, callbacks=[Val(), Probe(), es], initial_epoch=0,
                    logger=msgs.append, sync_stop=lambda f: f)
    assert hist["val_dice"] == [0.2, 0.3, 0.3, 0.3] and len(hist["loss"]) == 4       # stopped after epoch 3
    assert seen == list(range(12))
    assert order == ["begin", "eb0", "ee0", "eb1", "ee1", "eb2", "ee2", "eb3", "ee3", "end"]
    assert m.stop_training and len(msgs) == 4
    # initial_epoch resumes the numbering
    order.clear()
    fit_loop(m, ((i, None, None) for i in range(10)), steps_per_epoch=1, epochs=4, callbacks=[Val(), Probe()],
             initial_epoch=2, verbose=0)
    assert order == ["begin", "eb2", "ee2", "eb3", "ee3", "end"]


def test_output_bias_from_class_frequencies():
    """utils.set_bias_weights(_on_all_outputs) (mpunet/utils/utils.py:179-242): softmax(bias) is proportional to the
    class frequencies before the unit-norm scaling, only softmax output layers are accepted."""
    from multiplanarunet_b200.utils import utils as U

    class Layer:
        def __init__(self, act):
            self.activation = act
            self.w = [np.zeros((1, 1, 4, 3), np.float32), np.zeros(3, np.float32)]

        def get_weights(self):
            return [w.copy() for w in self.w]

        def set_weights(self, ws):
            self.w = ws

    def softmax(x):
        return x

    def relu(x):
        return x

    class Img:
        labels = np.array([0] * 70 + [1] * 20 + [2] * 10, dtype=np.uint8).reshape(10, 10)

    class M:
        layers = [Layer(relu), Layer(softmax)]
    msgs = []
    b = U.set_bias_weights_on_all_outputs(M, [Img, Img], {}, msgs.append)
    freq = np.array([0.7, 0.2, 0.1])
    raw = np.log(freq * np.exp(freq).sum())
    assert np.allclose(b, raw / np.linalg.norm(raw)) and np.allclose(M.layers[-1].w[-1], b.astype(np.float32))
    assert np.allclose(np.exp(raw) / np.exp(raw).sum(), freq)
    assert M.layers[0].w[-1].sum() == 0 and any("Estimating class counts from 2 images" in s for s in msgs)
    b2 = U.set_bias_weights(M.layers[-1], class_counts=[7, 2, 1], logger=msgs.append)
    assert np.allclose(b, b2)
    with pytest.raises(ValueError):
        U.set_bias_weights(M.layers[0], class_counts=[7, 2, 1])
