"""Mirror of critics.py's `critic` callable (critics.py:42-98) on torch tensors.

    cfg.critic = make_critic(trainer)                      # the WGAN critic ('critic' scope)
    logit, _, _ = cfg.critic(images=x, cfg=cfg)            # [B,1]
    value, _, _ = make_critic(trainer, value=True)(images=x, cfg=cfg, states=s)   # 'rl_value/critic' scope

`reuse` / `is_train` are accepted and ignored like in the reference (the weights are shared by
construction: they live in the Trainer)."""


def make_critic(trainer, value=False):
  net = trainer.value if value else trainer.critic

  def critic(images, cfg, states=None, is_train=None, reuse=False):
    if value:
      assert states is not None, "the value network takes the agent states (net.py:79-90)"
    else:
      assert states is None, "the WGAN critic sees images only (net.py:68-73)"
    c = net.forward(images.contiguous(), None if states is None else states.contiguous())
    return c.logit, None, None

  return critic
