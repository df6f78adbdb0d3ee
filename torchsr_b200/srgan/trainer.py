"""SRGAN training loop - mirror of the hot parts of torchsr/srgan/trainer.py.

Same class name, constructor signature, attribute and method names as the reference's SRGANTrainer for the parts
that sit on the hot path: model/loss/optimizer construction (reference :136-196), the pretrain step (:376-388), the
GAN step `_gan_loop` (:416-469), evaluation `_test` (:260-343) and the checkpoint format (:254-258, :321-327).
Differences, all math-identical:
  * the generator step runs the discriminator with its parameters frozen, so the weight gradients the reference
    computes, all-reduces and then discards at the next `discriminator.zero_grad()` are never computed;
  * data-parallel gradient averaging is done by torchsr_b200.dist (bucketed NCCL all-reduce of the flat gradient
    launched from inside backward) instead of torch DDP wrappers; BatchNorm statistics stay local to each GPU, as
    in the reference (no SyncBatchNorm).
"""
import os
import time
from argparse import Namespace
from math import log10
from typing import Optional

import torch
from torch import Tensor, optim

from .. import dist as tdist
from .. import losses
from ..optim import FusedAdam
from .discriminator import Discriminator
from .generator import Generator
from .loss import VGGLoss


class SRGANTrainer:
    PREFIX = 'srgan'

    def __init__(self, device, args: Namespace, train_loader, test_loader, train_len: int, test_len: int,
                 distributed: bool = False) -> None:
        self.amp = not args.disable_amp
        self.batch_size = args.batch_size
        self.best_psnr = -1.0
        self.device = torch.device(device)
        self.distributed = distributed
        self.epochs = args.epochs
        self.gan_checkpoint = args.gan_checkpoint
        self.local_rank = args.local_rank
        self.pre_epochs = args.pretrain_epochs
        self.psnr_checkpoint = args.psnr_checkpoint
        self.save_image = not args.skip_image_save
        self.test_loader, self.test_len = test_loader, test_len
        self.train_loader, self.train_len = train_loader, train_len
        self.world_size = args.world_size
        self.main_process = args.rank in [-1, 0]
        if self.device.type == 'cuda' and args.local_rank is not None and args.local_rank >= 0:
            torch.cuda.set_device(args.local_rank)
            self.device = torch.device('cuda', args.local_rank)
        if self.save_image and self.main_process and not os.path.exists('output'):
            os.makedirs('output')
        self._initialize_trainer()
        self._create_test_image()

    # ------------------------------------------------------------------ construction (reference :136-205)
    def _configure_modules(self) -> None:
        """Trainer-owned execution settings of the drop-in modules: gradients are handed to autograd as views of the
        plan's flat buffer (no copy) - valid here because every backward is preceded by zero_grad() - and a second
        CUDA stream carries the VGG content branch."""
        for m in (self.generator, self.discriminator):
            # gradients are handed to autograd as views of the plan's flat buffer (stable addresses, no copy); calls of
            # one module that feed one loss - D(real) and D(fake) - are summed by the last of their backward passes
            # (one add instead of one per parameter) and all-reduced in place once
            m._tsr["alias_grads"] = True
            m._tsr["merge_pending_grads"] = True
        cuda = self.device.type == 'cuda'
        prio = lambda name, default: int(os.environ.get(name, default))  # noqa: E731
        self._vgg_stream = torch.cuda.Stream(device=self.device, priority=prio("TSR_PRIO_VGG", -1)) if cuda else None
        self._capture_stream = torch.cuda.Stream(device=self.device, priority=prio("TSR_PRIO_MAIN", -1)) if cuda else None

    def _initialize_models(self) -> None:
        self.generator = Generator().to(self.device)
        self.discriminator = Discriminator().to(self.device)
        self._configure_modules()
        if self.distributed:
            # reference :143-157 wraps both in DistributedDataParallel; here: broadcast from rank 0 once, then the
            # modules all-reduce their own flat gradients (torchsr_b200/dist.py)
            tdist.attach(self.generator, broadcast_buffers=True)
            tdist.attach(self.discriminator, broadcast_buffers=False)

    def _initialize_loss(self) -> None:
        # same attribute names as the reference (:163-165); the criteria run on this repo's reduction kernels
        self.mse_loss = losses.MSELoss()
        self.bce_loss = losses.BCELoss()
        self.vgg_loss = VGGLoss().to(self.device)
        # upstream gradient of every scalar loss: passed explicitly so that backward() launches no fill kernel
        self._one = torch.ones((), dtype=torch.float32, device=self.device)

    def _initialize_optimizers(self) -> None:
        # Adam on this repo's kernels (optim.FusedAdam: one launch updates parameters, state and the packed bf16 weight
        # copies); tensor lr + device-side step counter so that a whole training step can be replayed as one CUDA graph
        # (graph_step) while StepLR keeps working (schedulers fill_() a tensor learning rate in place).
        mk = lambda params, **kw: FusedAdam(params, lr=torch.tensor(0.0001, device=self.device), betas=(0.9, 0.999), **kw)  # noqa: E731
        self.psnr_optimizer = mk(self.generator.parameters())
        # the classifier weight (80 % of the discriminator's parameters) is updated by a second launch on a side stream:
        # D(super_res) of the generator step starts on the updated conv weights meanwhile (joined at the end of the step)
        self.disc_optimizer = mk(self.discriminator.parameters(), late_numel=1 << 20)
        self.gen_optimizer = mk(self.generator.parameters())
        step = max(1, self.epochs // 8)   # reference :188 divides by zero for --epochs < 8
        self.disc_scheduler = optim.lr_scheduler.StepLR(self.disc_optimizer, step_size=step, gamma=0.6)
        self.gen_scheduler = optim.lr_scheduler.StepLR(self.gen_optimizer, step_size=step, gamma=0.6)

    def _initialize_trainer(self) -> None:
        self._initialize_models()
        self._initialize_loss()
        self._initialize_optimizers()

    def _create_test_image(self) -> None:
        self.test_image = None
        path = 'media/waterfalls-low-res.png'
        if self.save_image and os.path.exists(path):
            from PIL import Image
            from torchvision.transforms import ToTensor
            self.test_image = ToTensor()(Image.open(path).convert('RGB')).unsqueeze(0).to(self.device)

    def _log(self, statement: str) -> None:
        if self.main_process:
            print(statement)

    # ------------------------------------------------------------------ steps
    def _pretrain_step(self, low_res: Tensor, high_res: Tensor) -> Tensor:
        """One PSNR-phase step (reference :376-388): G forward, MSE, backward, Adam. The reference wraps this in fp16
        autocast + GradScaler; the kernels here compute in bf16 with fp32 accumulation, which needs no loss scaling."""
        low_res = low_res.to(self.device, non_blocking=True)
        high_res = high_res.to(self.device, non_blocking=True)
        self.psnr_optimizer.zero_grad()
        super_res = self.generator(low_res)
        loss = self.mse_loss(super_res, high_res)
        loss.backward(self._one)
        self.psnr_optimizer.step()
        tdist.join_buffers(self.generator)
        return loss.detach()

    def _gan_loop(self, low_res: Tensor, high_res: Tensor, step: int) -> Tensor:
        """One GAN step, statement for statement the reference's `_gan_loop` (:435-469)."""
        low_res = low_res.to(self.device, non_blocking=True)
        high_res = high_res.to(self.device, non_blocking=True)
        cur, vs = torch.cuda.current_stream(self.device), self._vgg_stream
        split_vgg = hasattr(self.vgg_loss, 'from_features')
        self.discriminator.zero_grad()
        # Two dependency chains, same arithmetic as the reference's sequential statements (:444-468):
        #   stream A (current): G forward -> D(real | fake) -> discriminator step -> D(super_res) -> G backward -> Adam
        #   stream V: everything of the VGG content loss - target features, features of super_res and the gradient of
        #             the content loss w.r.t. super_res. None of it depends on the discriminator, so it fills idle SMs
        #             during the discriminator step instead of sitting on the critical path of the generator step.
        vs.wait_stream(cur)
        if split_vgg:
            with torch.cuda.stream(vs):
                hr_feats = self.vgg_loss.target_features(high_res)
        super_res = self.generator(low_res)
        if split_vgg:
            vs.wait_stream(cur)
            super_res.record_stream(vs)
            with torch.cuda.stream(vs):
                # d(content)/d(super_res) through a detached leaf; added to the adversarial gradient below, which is
                # exactly what autograd does at the super_res node for gen_loss = content + 0.001 * adversarial
                sr_leaf = super_res.detach().requires_grad_(True)
                content_loss = self.vgg_loss.from_features(sr_leaf, hr_feats)
                (g_content,) = torch.autograd.grad(content_loss, sr_leaf, grad_outputs=self._one)
                content_loss = content_loss.detach()
                done_content = vs.record_event()
        # reference :446-447 calls D(high_res) and D(super_res.detach()) separately: separate BatchNorm batch statistics
        # and two running-statistics updates, real first. forward_pair runs both batches through ONE pass of the
        # discriminator's kernels with exactly those semantics (statistics per half, engine.py bn_groups).
        p_real, p_fake = self.discriminator.forward_pair(high_res, super_res.detach())
        # BCE(p_real, 1) + BCE(p_fake, 0) (:446-448) in one reduction launch, constant labels
        disc_loss = losses.bce(p_real, 1.0, p_fake, 0.0)
        with self.disc_optimizer.late_in_backward(self.discriminator):
            disc_loss.backward(self._one)
        self.disc_optimizer.step()

        self.generator.zero_grad()
        with tdist.frozen(self.discriminator):
            adversarial_loss = losses.bce(self.discriminator(super_res), 1.0, scale=0.001)     # 0.001 * BCE (:456-457)
        if split_vgg:
            (g_adv,) = torch.autograd.grad(adversarial_loss, super_res, grad_outputs=self._one)
            cur.wait_event(done_content)
            g_content.record_stream(cur)
            content_loss.record_stream(cur)
            super_res.backward(losses.add(g_content, g_adv))
            gen_loss = losses.add(content_loss, adversarial_loss)
        else:
            content_loss = self.vgg_loss(super_res, high_res.detach())
            gen_loss = losses.total(content_loss, adversarial_loss)
            gen_loss.backward(self._one)
        self.gen_optimizer.step()
        self.disc_optimizer.join()
        tdist.join_buffers(self.generator)
        return gen_loss.detach()

    # ------------------------------------------------------------------ whole-step CUDA graph
    def graph_step(self, low_res: Tensor, high_res: Tensor, step: int = 0, kind: str = 'gan') -> Tensor:
        """`_gan_loop` (kind='gan') or `_pretrain_step` (kind='psnr') replayed as ONE CUDA graph. The first call for a
        batch shape runs the step eagerly (that IS the step for this batch: builds plans, launch lists and optimizer
        tables, sizes the plan pools) and then records the same step into a graph without executing it; every later
        call copies the batch into static input buffers and replays. Results are identical to the eager step; from
        the second call on the returned loss tensor is a static buffer that the next call overwrites."""
        key = (kind, tuple(low_res.shape), tuple(high_res.shape))
        g = self._graphs.get(key) if hasattr(self, '_graphs') else None
        if g is None:
            if not hasattr(self, '_graphs'):
                self._graphs = {}
            fn = (lambda a, b: self._gan_loop(a, b, step)) if kind == 'gan' else self._pretrain_step
            s_lr = torch.empty(low_res.shape, dtype=torch.float32, device=self.device)
            s_hr = torch.empty(high_res.shape, dtype=torch.float32, device=self.device)
            s_lr.copy_(low_res, non_blocking=True)
            s_hr.copy_(high_res, non_blocking=True)
            first_loss = fn(s_lr, s_hr).clone()
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=self._capture_stream):     # capture only: nothing executes here
                loss = fn(s_lr, s_hr)
            self._graphs[key] = (graph, s_lr, s_hr, loss)
            return first_loss
        graph, s_lr, s_hr, loss = g
        s_lr.copy_(low_res, non_blocking=True)
        s_hr.copy_(high_res, non_blocking=True)
        graph.replay()
        return loss

    def _train_step(self, kind: str, low_res: Tensor, high_res: Tensor, step: int) -> Tensor:
        """The step the epoch loops call: the whole-step CUDA graph for full batches on a GPU (TORCHSR_GRAPH_STEP=0
        selects the eager call), the eager step for a ragged batch."""
        if (self.device.type == 'cuda' and low_res.size(0) == self.batch_size
                and os.environ.get('TORCHSR_GRAPH_STEP', '1') != '0'):
            return self.graph_step(low_res, high_res, step, kind)
        return self._gan_loop(low_res, high_res, step) if kind == 'gan' else self._pretrain_step(low_res, high_res)

    # ------------------------------------------------------------------ evaluation / checkpoints
    def _model_state(self, epoch: int, phase: str) -> dict:
        return {"epoch": epoch, "phase": phase, "state": self.generator.state_dict()}

    def _test(self, epoch: int, phase: str, step: int) -> float:
        """Eval-mode, no-grad pass over the test loader: PSNR = 10 log10(1 / mse) per batch (reference :282-303)."""
        self.generator.eval()
        psnr, n = 0.0, 0
        with torch.no_grad():
            for low_res, _, high_res in self.test_loader:
                super_res = self.generator(low_res.to(self.device))
                psnr += 10 * log10(1 / ((super_res - high_res.to(self.device)) ** 2).mean().item())
                n += 1
        psnr = psnr / max(n, 1)
        self._log(f'PSNR: {round(psnr, 3)}')
        if self.main_process:
            if psnr > self.best_psnr:
                self.best_psnr = psnr
                torch.save(self._model_state(epoch, phase), f'{phase}-best.pth')
            torch.save(self._model_state(epoch, phase), f'{phase}-latest.pth')
            if self.save_image and self.test_image is not None:
                from torchvision import utils
                utils.save_image(self.generator(self.test_image), f'output/SR_epoch{epoch}.png', padding=5)
        self.generator.train()
        return psnr

    def _load_checkpoint(self, path: Optional[str]):
        if path and os.path.exists(path):
            return torch.load(path, map_location=self.device)
        return None

    def _restore(self, *paths) -> Optional[dict]:
        """Loads the generator weights from the first checkpoint that exists and returns it (None if none does).
        Accepts the trainer's {"epoch", "phase", "state"} payload and bare (optionally `module.`-prefixed) state dicts."""
        for p in paths:
            ck = self._load_checkpoint(p)
            if ck is not None:
                state = ck["state"] if "state" in ck else ck
                state = {k[len('module.'):] if k.startswith('module.') else k: v for k, v in state.items()}
                self.generator.load_state_dict(state)
                return ck if "state" in ck else {"state": state}
        return None

    def _check_device(self) -> None:
        """Epoch boundary (already synchronised): raise if a kernel's pipeline watchdog fired during the epoch instead
        of training on with a garbage tile (csrc/ptx.cuh mbar_wait)."""
        if self.device.type == 'cuda':
            from .. import ops
            ops.check_watchdog()

    def _pretrain(self) -> None:
        self.best_psnr = -1.0
        # reference :357-364: explicit --psnr-checkpoint, else <model>-psnr-latest.pth; training resumes at its epoch
        ck = self._restore(self.psnr_checkpoint) if self.psnr_checkpoint else self._restore(f'{self.PREFIX}-psnr-latest.pth')
        first = int(ck.get("epoch", 1)) if ck else 1
        step = 0
        for epoch in range(first, self.pre_epochs + 1):
            self._log(f'Starting epoch {epoch} out of {self.pre_epochs}')
            t0 = time.time()
            seen = 0
            for low_res, high_res in self.train_loader:
                self._train_step('psnr', low_res, high_res, step)
                seen += low_res.size(0)
                step += 1
            if self.device.type == 'cuda':
                torch.cuda.synchronize()
            self._check_device()
            self._log(f'Throughput: {round(seen * max(self.world_size, 1) / (time.time() - t0), 3)} images/sec')
            self._test(epoch, f'{self.PREFIX}-psnr', step)

    def _gan_train(self) -> None:
        self.best_psnr = -1.0
        # reference :483-499: a GAN checkpoint (explicit or <model>-gan-latest.pth) restores weights AND the epoch to
        # resume at; without one the PSNR-phase weights seed the generator and the GAN phase starts at epoch 1
        ck = self._restore(self.gan_checkpoint) if self.gan_checkpoint else self._restore(f'{self.PREFIX}-gan-latest.pth')
        first = int(ck.get("epoch", 1)) if ck else 1
        if ck is None:
            self._restore(f'{self.PREFIX}-psnr-latest.pth')
        step = 0
        for epoch in range(first, self.epochs + 1):
            self._log(f'Starting epoch {epoch} out of {self.epochs}')
            t0 = time.time()
            seen = 0
            for low_res, high_res in self.train_loader:
                self._train_step('gan', low_res, high_res, step)
                seen += low_res.size(0)
                step += 1
            self.disc_scheduler.step()
            self.gen_scheduler.step()
            if self.device.type == 'cuda':
                torch.cuda.synchronize()
            self._check_device()
            self._log(f'Throughput: {round(seen * max(self.world_size, 1) / (time.time() - t0), 3)} images/sec')
            self._test(epoch, f'{self.PREFIX}-gan', step)

    def train(self) -> None:
        self._pretrain()
        if self.device.type == 'cuda':
            torch.cuda.empty_cache()
        self._gan_train()
