"""`specforge train` with the EAGLE3 step served by libspecforge_b200 — the drop-in entry point.

    python -m specforge_b200.train train -c examples/configs/qwen3-8b-eagle3-offline.yaml [overrides ...]

Needs the reference package importable (a SpecForge checkout / install on sys.path; in this repo's tests:
`baseline/_ref`).  Everything the reference does around the step is the reference's own code, unchanged: config loading
and validation, launch planning, the offline manifest reader / feature store / collator, `Trainer`, `TrainerController`,
`TrainerCore.train_step`, checkpoint manager, trackers, export.  What is replaced, through the reference's own seams
(the stock CLI hard-wires the builtin catalog, cli.py:246, hence this wrapper):

  seam 4  AlgorithmRegistry with an "eagle3" registration whose providers build OUR pieces
          (algorithms/registry.py:78-91, algorithms/eagle3/providers.py:117-236)
  seam 1  ModelProvider.build_draft           -> B200Eagle3DraftModel (same state-dict keys / helpers), engine bound to
                                                 (training.batch_size, data.max_length, training.ttt_length)
          ModelProvider.build_training_model  -> the composite exposing `.draft_model` (+ the attributes the resume contract
                                                 reads, providers.py:62-78) and the reference's own frozen `TargetHead`
  seam 2  StepProvider.build                  -> B200Eagle3TrainStrategy (providers.py:45-52)
  seam 3  Trainer's backend constructor (training/trainer.py:421) -> B200TrainingBackend, and the configured optimizer
          factory's `BF16Optimizer` (training/assembly.py:246-274) -> B200BF16Optimizer

`b200_patches()` is the context manager that swaps those two names for the duration of a run; `main()` mirrors
`specforge.cli.main`'s `train` branch (cli.py:241-267) with `resolve_run(cfg, registry=b200_registry())`.
"""
from __future__ import annotations

import contextlib
import dataclasses
import os
import sys
from typing import List, Optional

import torch
import torch.nn as nn

from .backend import B200TrainingBackend
from .draft import B200Eagle3DraftModel
from .optimizer import B200BF16Optimizer
from .strategy import B200Eagle3TrainStrategy

ALGORITHM_NAME = "eagle3"


class B200Eagle3Model(nn.Module):
    """What `build_training_model` returns in place of OnlineEagle3Model (algorithms/eagle3/model.py:244-442): the trainable
    composite exposing `.draft_model` (trainer.py:205,425) plus the settings the EAGLE3 resume contract persists."""

    def __init__(self, draft_model: B200Eagle3DraftModel, *, length: int, attention_backend: str, lk_loss_type, kl_scale: float,
                 kl_decay: float):
        super().__init__()
        self.draft_model = draft_model
        self.length = length
        self.attention_backend = attention_backend
        self.lk_loss_type = lk_loss_type
        self.kl_scale = kl_scale
        self.kl_decay = kl_decay

    def forward(self, *args, **kwargs):
        raise RuntimeError("the B200 EAGLE3 model is driven through B200Eagle3TrainStrategy.forward_loss (no eager forward)")


# ---- providers (same signatures as specforge/algorithms/eagle3/providers.py) -------------------------------------------------
def build_step(wrapped_model, *, target_head=None, **options):
    return B200Eagle3TrainStrategy(wrapped_model, target_head=target_head, **options)


def build_draft(config, draft_config):
    """algorithms/model_providers.py:91-112 (build_eagle3_draft) with our module: warm start, vocab mapping, frozen embedding."""
    from specforge.algorithms import model_providers as ref
    t = config.training
    draft = B200Eagle3DraftModel(draft_config, attention_backend=t.attention_backend)
    draft.bind_engine(batch=t.batch_size, seq_len=config.data.max_length, ttt_length=t.ttt_length, device=_local_device(),
                      lk_loss_type=t.lk_loss_type, kl_scale=t.kl_scale, kl_decay=t.kl_decay)
    ref._warm_start(config, draft, draft_config, allow_missing_embedding=True)
    ref._load_vocab_mapping(config, draft)
    if config.model.load_target_embedding:
        draft.load_embedding(config.model.target_model_path, embedding_key=config.model.embedding_key)
    draft.freeze_embedding()
    return draft


def build_training_model(config, draft_model, draft_config, target_config, tokenizer):
    """algorithms/model_providers.py:255-288 (build_eagle3_model): the composite + the reference's frozen TargetHead."""
    from specforge.algorithms.model_providers import AlgorithmModelParts
    t = config.training
    model = B200Eagle3Model(draft_model, length=t.ttt_length, attention_backend=t.attention_backend, lk_loss_type=t.lk_loss_type,
                            kl_scale=t.kl_scale, kl_decay=t.kl_decay)
    needs_target_head = config.mode == "offline" or (config.deployment.mode == "disaggregated" and t.role == "consumer")
    target_head = None
    if needs_target_head:
        from specforge.modeling.target.target_head import TargetHead
        target_head = TargetHead.from_pretrained(config.model.target_model_path, lm_head_key=config.model.lm_head_key,
                                                 cache_dir=config.model.cache_dir, trust_remote_code=config.model.trust_remote_code)
    return AlgorithmModelParts(model=model, target_head=target_head)


def _local_device() -> torch.device:
    return torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))


def eagle3_registration():
    """The reference's EAGLE3 registration (contracts, data providers, capture layouts untouched) with the step and model
    factories above."""
    from specforge.algorithms.common.providers import make_registration
    from specforge.algorithms.eagle3 import providers as ref
    prov = ref.algorithm_providers()
    step = dataclasses.replace(prov.step, build=build_step)
    model = dataclasses.replace(prov.model, build_draft=build_draft, build_training_model=build_training_model)
    return make_registration(ref.algorithm_spec(), dataclasses.replace(prov, step=step, model=model))


def b200_registry():
    """AlgorithmRegistry((ours, peagle, dflash, domino, dspark)): `training.strategy: eagle3` resolves to the CUDA path, every
    other algorithm to the reference's own (algorithms/builtin.py:13-16)."""
    from specforge.algorithms.builtin import builtin_algorithm_registry
    from specforge.algorithms.registry import AlgorithmRegistry
    others = [r for r in builtin_algorithm_registry().registrations if r.name != ALGORITHM_NAME]
    return AlgorithmRegistry((eagle3_registration(), *others))


@contextlib.contextmanager
def b200_patches():
    """Seam 3: `Trainer` names its backend class and the optimizer factory names `BF16Optimizer` directly
    (training/trainer.py:421, training/assembly.py:261); both are module attributes, swapped here for the run."""
    import specforge.optimizer as ref_opt
    import specforge.training.trainer as ref_trainer
    saved = (ref_trainer.FSDPTrainingBackend, ref_opt.BF16Optimizer)
    ref_trainer.FSDPTrainingBackend = B200TrainingBackend
    ref_opt.BF16Optimizer = B200BF16Optimizer
    try:
        yield
    finally:
        ref_trainer.FSDPTrainingBackend, ref_opt.BF16Optimizer = saved


def main(argv: Optional[List[str]] = None) -> int:
    """`specforge train` (cli.py:241-267) with the B200 registry; other sub-commands are the reference's own."""
    import argparse
    from specforge import cli as ref_cli
    from specforge.application import bind_run, resolve_run
    from specforge.config import load_config
    from specforge.launch_plan import build_launch_plan, run_commands
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] != "train":
        return ref_cli.main(argv)
    ap = argparse.ArgumentParser(prog="specforge-b200 train")
    ap.add_argument("-c", "--config", required=True)
    ap.add_argument("--role", choices=("auto", "all", "producer", "consumer", "both"), default="auto")
    ap.add_argument("--node-rank", type=int, default=None)
    ap.add_argument("--plan", action="store_true")
    ap.add_argument("overrides", nargs="*")
    args = ap.parse_args(argv[1:])
    cfg = load_config(args.config, args.overrides)
    resolved = resolve_run(cfg, registry=b200_registry())
    plan = build_launch_plan(resolved.config, algorithm=resolved.algorithm, config_path=args.config, overrides=args.overrides,
                             requested_role=args.role, node_rank=args.node_rank)
    if args.plan:
        print(plan.render())
        return 0
    if plan.kind == "worker":
        os.environ.update(plan.worker_env)
        role_config = ref_cli._config_for_role(resolved.config, plan.role)
        with b200_patches():
            try:
                with ref_cli._worker_signal_unwind():
                    ref_cli._train(bind_run(role_config, resolved.algorithm))
            except ref_cli._WorkerTermination as received:
                return 128 + received.signum
        return 0
    # a multi-process plan re-invokes this module per worker (the plan's commands name the reference CLI module)
    return run_commands(_retarget(plan))


def _retarget(plan):
    """Launch plans spell worker commands as `python -m specforge.cli train ...`; point them at this module instead."""
    def fix(cmd):
        return [("specforge_b200.train" if a in ("specforge.cli", "specforge") else a) for a in cmd]
    try:
        return dataclasses.replace(plan, commands=tuple(dataclasses.replace(c, argv=tuple(fix(c.argv))) for c in plan.commands))
    except Exception:   # a plan shape this wrapper does not know: fail loudly rather than silently train on the stock path
        raise RuntimeError("launch plan layout not recognised; start one worker per GPU with torchrun -m specforge_b200.train")


if __name__ == "__main__":
    raise SystemExit(main())
