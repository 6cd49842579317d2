"""Training driver -- the counterpart of UNet/train.py: same `train_model(...)` signature, same command-line flags,
same side effects (`<out>/checkpoint/ckpt`, `<out>/test_loss.csv`, `<out>/tensorboard-<time>/{train,test}`, the
per-step `Train Epoch {e}: Batch {s}/{n}: Loss {} Accuracy = {}` lines).

    python -m unetb200.train --train_database DIR --test_database DIR --output_dir OUT [--batch_size 4] ...
    torchrun --nproc-per-node G -m unetb200.train ...        # one process per GPU instead of MirroredStrategy

Reference behaviour kept (file:line under /root/reference/UNet/train.py): `batch_size` is per GPU and the global batch
is batch_size x replicas (:61); the first epoch runs min(1000, test_every_n_steps) steps at learning_rate / 10 (:126-129);
an epoch of n steps executes steps 0..n inclusive (:136-138); the test epoch runs count / batch_size (+1)
steps (:100, :155); test loss = mean of the per-step SUM-reduced losses (:159-161); a checkpoint is written whenever
the latest test loss is the running minimum (:181-184); early stopping counts epochs since the first epoch within
1e-4 of the best loss (:187-199).
What is different underneath: batches travel as raw pixels + uint8 labels from pinned memory and are normalised on
the GPU; the step is unetb200.model.UNet.train_step; the per-step metric read-back is one step behind the GPU so the
device never waits for the host print.
"""
from __future__ import annotations

import argparse
import datetime
import os
import time

import numpy as np


class StepPipeline:
    """Training steps fed from PINNED HOST batches without stalling the device: the upload of batch i + 1 runs on a copy stream while
    step i computes, and the loss of step i is read back while step i + 1 runs (the reference hides both behind
    `dataset.prefetch` and TensorFlow's asynchronous metrics, UNet/train.py:85, :140-146).  Every batch is still copied host -> device
    and every step's loss still reaches the host; only their latency leaves the critical path.  One batch of look-ahead, two device
    slots (a slot is rewritten only after the step that consumed it has finished).

        pipe = StepPipeline(unet)
        for images, labels in batches:            # pinned float32 [B, C, H, W] images, uint8 [B, H, W] (or one-hot int32) labels
            loss = pipe.feed(images, labels)      # python float of the step BEFORE the one just launched, None at the start
        losses_tail = pipe.flush()                # runs the last uploaded batch; returns the losses not yet handed out
    """

    def __init__(self, unet):
        import torch
        self.torch = torch
        self.m = unet
        self.dev = unet.device
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.slots = [None, None]
        self.step_done = [None, None]
        self.k = 0
        self.ready = None           # (x, labels, upload-done event, slot) of the batch waiting for its step
        self.pending = None         # (pinned loss, event) of the last launched step

    def _upload(self, images, labels):
        torch = self.torch
        k = self.k
        self.k ^= 1
        slot = self.slots[k]
        if slot is None or slot[0].shape != images.shape or slot[0].dtype != images.dtype or slot[1].shape != labels.shape \
                or slot[1].dtype != labels.dtype:
            slot = (torch.empty(images.shape, dtype=images.dtype, device=self.dev), torch.empty(labels.shape, dtype=labels.dtype, device=self.dev))
            self.slots[k] = slot
        if self.step_done[k] is not None:
            self.copy_stream.wait_event(self.step_done[k])
        with torch.cuda.stream(self.copy_stream):
            slot[0].copy_(images, non_blocking=True)
            slot[1].copy_(labels, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return slot[0], slot[1], ev, k

    def _run(self, item):
        torch = self.torch
        x, lab, ev, k = item
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(ev)
        self.m.train_step(x, lab)
        done = torch.cuda.Event()
        done.record(cur)
        self.step_done[k] = done
        prev = self.pending
        vals = torch.empty(2, dtype=torch.float32, pin_memory=True)
        vals.copy_(self.m.metrics[:2], non_blocking=True)          # ordered on the step's stream, before the next step overwrites it
        got = torch.cuda.Event()
        got.record(cur)
        self.pending = (vals, got)
        if prev is None:
            return None
        prev[1].synchronize()                                      # the host runs at most one step ahead of the device
        return float(prev[0][0])

    def feed(self, images, labels):
        nxt = self._upload(images, labels)
        out = self._run(self.ready) if self.ready is not None else None
        self.ready = nxt
        return out

    def flush(self):
        out = []
        if self.ready is not None:
            v = self._run(self.ready)
            self.ready = None
            if v is not None:
                out.append(v)
        if self.pending is not None:
            self.pending[1].synchronize()
            out.append(float(self.pending[0][0]))
            self.pending = None
        return out


class _Mean:
    """stand-in for tf.keras.metrics.Mean / CategoricalAccuracy as UNet/train.py:104-107 uses them: the step hands it
    a 0-d device tensor; values are accumulated on the device and read only by result()"""

    def __init__(self, name):
        self.name = name
        self.reset_states()

    def update_state(self, value):
        self._sum = value.detach().clone() if self._sum is None else self._sum + value
        self._n += 1

    def result(self):
        return float(self._sum.item()) / self._n if self._n else 0.0

    def reset_states(self):
        self._sum = None
        self._n = 0


class _Summary:
    def __init__(self, log_dir):
        os.makedirs(log_dir, exist_ok=True)
        try:
            from torch.utils.tensorboard import SummaryWriter
            self._w = SummaryWriter(log_dir)
        except Exception as e:          # tensorboard is optional tooling, not part of the hot path
            print('tensorboard writer unavailable ({}); scalars are not logged'.format(e))
            self._w = None

    def scalar(self, tag, value, step):
        if self._w is not None:
            self._w.add_scalar(tag, value, step)

    def close(self):
        if self._w is not None:
            self._w.close()


def select_best_epoch(test_loss, tolerance=1e-4):
    """first epoch whose loss is within `tolerance` of the minimum (UNet/train.py:187-196)"""
    tl = np.asarray(test_loss, dtype=np.float64)
    err = np.abs(tl - np.min(tl))
    err[err < tolerance] = 0
    return int(np.where(err == 0)[0][0])


def train_model(output_folder, batch_size, reader_count, train_lmdb_filepath, test_lmdb_filepath, use_augmentation,
                number_classes, balance_classes, learning_rate, test_every_n_steps, early_stopping_count, max_epochs=None):
    import torch
    from . import imagereader, model
    from .dist import DataParallel

    print('batch_size = {}'.format(batch_size))
    print('number_classes = {}'.format(number_classes))
    print('learning_rate = {}'.format(learning_rate))
    print('test_every_n_steps = {}'.format(test_every_n_steps))
    print('balance_classes = {}'.format(balance_classes))
    print('use_augmentation = {}'.format(use_augmentation))
    print('train_database = {}'.format(train_lmdb_filepath))
    print('test_database = {}'.format(test_lmdb_filepath))
    print('output folder = {}'.format(output_folder))
    print('early_stopping count = {}'.format(early_stopping_count))
    print('reader_count = {}'.format(reader_count))

    os.makedirs(output_folder, exist_ok=True)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    strategy = DataParallel() if world > 1 else None
    rank = strategy.rank if strategy else 0
    replicas = strategy.num_replicas_in_sync if strategy else 1
    if strategy is None:
        torch.cuda.set_device(0)
    is_chief = rank == 0
    global_batch_size = batch_size * replicas          # UNet/train.py:61

    print('Setting up test image reader')
    test_reader = imagereader.ImageReader(test_lmdb_filepath, use_augmentation=False, shuffle=False, num_workers=reader_count,
                                          balance_classes=False, number_classes=number_classes, rank=rank, world_size=replicas)
    print('Test Reader has {} images'.format(test_reader.get_image_count()))
    print('Setting up training image reader')
    train_reader = imagereader.ImageReader(train_lmdb_filepath, use_augmentation=use_augmentation, shuffle=True, num_workers=reader_count,
                                           balance_classes=balance_classes, number_classes=number_classes,
                                           seed=None if replicas == 1 else 1000003 * (rank + 1) + int(time.time()), rank=rank, world_size=replicas)
    print('Train Reader has {} images'.format(train_reader.get_image_count()))

    try:
        print('Starting Readers')
        train_reader.startup()
        print('  train_reader online')
        test_reader.startup()
        print('  test_reader online')

        print('Creating model')
        number_channels = train_reader.get_image_size()[2]
        unet_model = model.UNet(number_classes, global_batch_size, number_channels, learning_rate, dist=strategy)
        if strategy is not None:
            strategy.broadcast_params(unet_model)
        dev = unet_model.device

        train_epoch_size = test_every_n_steps
        test_epoch_size = test_reader.get_image_count() / batch_size          # per-GPU batch, as the reference (Q4)
        test_loss = list()

        train_loss_metric, train_acc_metric = _Mean('train_loss'), _Mean('train_accuracy')
        test_loss_metric, test_acc_metric = _Mean('test_loss'), _Mean('test_accuracy')

        current_time = datetime.datetime.now().strftime("%Y%m%dT%H%M%S")
        train_summary = _Summary(os.path.join(output_folder, 'tensorboard-' + current_time, 'train')) if is_chief else None
        test_summary = _Summary(os.path.join(output_folder, 'tensorboard-' + current_time, 'test')) if is_chief else None

        def device_batch(reader):
            return reader.device_batch(batch_size, unet_model)      # upload, z-score (test reader: on the step's stream)

        class _Prefetch:
            """The next training batch is uploaded, augmented and z-scored on a side stream while the current step runs (the
            reference's dataset.prefetch, UNet/train.py:85).  Two slots: a slot is rewritten only after the step that consumed
            it has finished on the device."""

            def __init__(self, reader):
                self.reader = reader
                self.stream = torch.cuda.Stream(device=dev)
                self.slot = 0
                self.step_done = [None, None]
                self.next = None

            def _issue(self):
                k = self.slot
                self.slot ^= 1
                if self.step_done[k] is not None:
                    self.stream.wait_event(self.step_done[k])
                with torch.cuda.stream(self.stream):
                    x, lab = self.reader.device_batch(batch_size, unet_model, slot=k)
                    ev = torch.cuda.Event()
                    ev.record(self.stream)
                lab.record_stream(torch.cuda.current_stream(dev))         # consumed by the step on the main stream
                return x, lab, ev, k

            def get(self):
                if self.next is None:
                    self.next = self._issue()
                x, lab, ev, k = self.next
                torch.cuda.current_stream(dev).wait_event(ev)
                self.next = self._issue()
                return x, lab, k

            def consumed(self, k):
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
                self.step_done[k] = ev

        prefetch = _Prefetch(train_reader)

        epoch = 0
        print('Running Network')
        while True:
            print('---- Epoch: {} ----'.format(epoch))
            if epoch == 0:
                cur_train_epoch_size = min(1000, train_epoch_size)
                print('Performing Adam Optimizer learning rate warmup for {} steps'.format(cur_train_epoch_size))
                unet_model.set_learning_rate(learning_rate / 10)
            else:
                cur_train_epoch_size = train_epoch_size
                unet_model.set_learning_rate(learning_rate)

            start_time = time.time()
            pending = None          # (step, pinned [loss, accuracy], event) of the previous step: printed while the next one runs

            def flush(p):
                step_p, vals, ev = p
                ev.synchronize()
                lv, av = float(vals[0]), float(vals[1])
                print('Train Epoch {}: Batch {}/{}: Loss {} Accuracy = {}'.format(epoch, step_p, train_epoch_size, lv, av))
                if train_summary is not None:
                    train_summary.scalar('loss', lv, int(epoch * train_epoch_size + step_p))
                    train_summary.scalar('accuracy', av, int(epoch * train_epoch_size + step_p))

            step = 0
            while step <= cur_train_epoch_size:          # steps 0..n inclusive (UNet/train.py:136-138)
                x, lab, slot = prefetch.get()
                unet_model.dist_train_step(strategy, (x, lab, train_loss_metric, train_acc_metric))
                prefetch.consumed(slot)
                if pending is not None:
                    flush(pending)
                vals = torch.empty(2, dtype=torch.float32, pin_memory=True)
                vals.copy_(unet_model.metrics, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                pending = (step, vals, ev)
                train_loss_metric.reset_states()
                train_acc_metric.reset_states()
                step += 1
            if pending is not None:
                flush(pending)

            epoch_test_loss = list()
            step = 0
            while step <= test_epoch_size:               # UNet/train.py:154-161
                x, lab = device_batch(test_reader)
                loss_value = unet_model.dist_test_step(strategy, (x, lab, test_loss_metric, test_acc_metric))
                epoch_test_loss.append(loss_value)
                step += 1
            test_loss.append(float(np.mean([float(v.item()) for v in epoch_test_loss])))

            tl, ta = test_loss_metric.result(), test_acc_metric.result()
            print('Test Epoch: {}: Loss = {} Accuracy = {}'.format(epoch, tl, ta))
            if test_summary is not None:
                test_summary.scalar('loss', tl, int((epoch + 1) * train_epoch_size))
                test_summary.scalar('accuracy', ta, int((epoch + 1) * train_epoch_size))
            test_loss_metric.reset_states()
            test_acc_metric.reset_states()

            if is_chief:
                with open(os.path.join(output_folder, 'test_loss.csv'), 'w') as csvfile:
                    for v in test_loss:
                        csvfile.write(str(v))
                        csvfile.write('\n')
            print('Epoch took: {} s'.format(time.time() - start_time))

            if (len(test_loss) - 1) == int(np.argmin(test_loss)):
                print('Test loss improved: {}, saving checkpoint'.format(np.min(test_loss)))
                import contextlib
                # ON_READ / MEAN aggregation of the moving statistics for the checkpoint only: replicas keep their own
                with (strategy.moving_stats_averaged(unet_model) if strategy is not None else contextlib.nullcontext()):
                    if is_chief:
                        os.makedirs(os.path.join(output_folder, 'checkpoint'), exist_ok=True)
                        unet_model.save_checkpoint(os.path.join(output_folder, 'checkpoint', "ckpt"))

            print('Best Current Epoch Selection:')
            print('Test Loss:')
            print(test_loss)
            best_epoch = select_best_epoch(test_loss)
            print('Best epoch: {}'.format(best_epoch))
            if len(test_loss) - best_epoch > early_stopping_count:
                break
            epoch = epoch + 1
            if max_epochs is not None and epoch >= max_epochs:
                break
        for s in (train_summary, test_summary):
            if s is not None:
                s.close()
        return test_loss
    finally:
        print('Shutting down train_reader')
        train_reader.shutdown()
        print('Shutting down test_reader')
        test_reader.shutdown()
        if strategy is not None:
            strategy.shutdown(locals().get("unet_model"))


def main(argv=None):
    parser = argparse.ArgumentParser(prog='train_unet', description='Script which trains a unet model')
    parser.add_argument('--train_database', dest='train_database_filepath', type=str, help='lmdb database to use for (Required)', required=True)
    parser.add_argument('--test_database', dest='test_database_filepath', type=str, help='lmdb database to use for testing (Required)', required=True)
    parser.add_argument('--output_dir', dest='output_folder', type=str, help='Folder where outputs will be saved (Required)', required=True)
    parser.add_argument('--batch_size', dest='batch_size', type=int, help='training batch size', default=4)
    parser.add_argument('--number_classes', dest='number_classes', type=int, default=2)
    parser.add_argument('--learning_rate', dest='learning_rate', type=float, default=3e-4)
    parser.add_argument('--test_every_n_steps', dest='test_every_n_steps', type=int, help='number of gradient update steps to take between test epochs', default=1000)
    parser.add_argument('--balance_classes', dest='balance_classes', type=int, help='whether to balance classes [0 = false, 1 = true]', default=0)
    parser.add_argument('--use_augmentation', dest='use_augmentation', type=int, help='whether to use data augmentation [0 = false, 1 = true]', default=1)
    parser.add_argument('--early_stopping', dest='early_stopping_count', type=int, help='Perform early stopping when the test loss does not improve for N epochs.', default=10)
    parser.add_argument('--reader_count', dest='reader_count', type=int, help='accepted for drop-in compatibility and ignored: the reference forks this many reader processes per gpu '
                             '(UNet/imagereader.py:175-186); here records are decoded in-process into pinned buffers (2800 images/s per process) and '
                             'augmentation runs on the GPU one batch ahead of the step', default=1)
    args = parser.parse_args(argv)
    train_model(args.output_folder, args.batch_size, args.reader_count, args.train_database_filepath, args.test_database_filepath,
                args.use_augmentation, args.number_classes, args.balance_classes, args.learning_rate, args.test_every_n_steps,
                args.early_stopping_count)


if __name__ == "__main__":
    main()
