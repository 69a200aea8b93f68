"""Mirror of agent.py's `agent_generator` callable (agent.py:41-260) on torch tensors.

The reference builds a TF sub-graph inside `tf.variable_scope('generator')`; its weights live in the
graph.  Here the weights live in a `Trainer` (exposure_b200.trainer), so the callable is bound to one:

    cfg.generator = make_agent_generator(trainer)
    (net, new_states, surrogate, penalty), debug_info, debugger = cfg.generator(
        inp=[img, z, states], is_train=1, progress=0.3, cfg=cfg)

Same argument meaning and return structure as the reference:
  inp       [fake_input [B,64,64,3], z [B, cfg.z_dim] (only z[:, 0] is used, agent.py:47), states [B,11]]
  is_train  1: sample the action from the pdf, 0: argmax                      (agent.py:113-116)
  progress  training progress in [0,1] (entropy-penalty schedule, agent.py:240-244)
  high_res  optional [B,H,W,3]: the selected filter is also applied to it    (agent.py:61-63,257-260)
Forward only (no gradients): training goes through Trainer.generator_step, which is the explicit
forward + backward schedule of the same graph."""
import torch


def make_agent_generator(trainer):
  policy = trainer.policy

  def agent_generator(inp, is_train, progress, cfg, high_res=None, alex_in=None):
    assert alex_in is None, "alex_in is unused by the reference as well (agent.py:41)"
    net, z, states = inp
    B = net.shape[0]
    dev = net.device
    noise = z[:, 0].contiguous()
    drop_f, drop_s = trainer.draw(B)[1:3]              # tf.nn.dropout has no is_train switch (agent.py:36): always on
    prog = torch.full((1,), float(progress), device=dev)
    c = policy.forward(net.contiguous(), states.contiguous(), noise, drop_f, drop_s, int(is_train), prog, cfg,
                       high_res=high_res)
    debug_info = {
        "selected_filter_id": c.ids,                                   # agent.py:117-118
        "filter_parameters": c.params,                                 # parameters of the SELECTED filter
        "pdf": getattr(c, "pdf", None),
        "penalty": c.penalty,
    }

    def debugger(*args, **kwargs):                                     # agent.py:141-202 (visualisation): out of scope
      raise NotImplementedError("the visual debugger of agent.py:141-202 is outside the hot path (DESIGN.md section 10)")

    if high_res is None:
      return (c.out, c.new_states, c.surrogate.view(B, 1), c.penalty.view(B, 1)), debug_info, debugger
    return (c.out, c.new_states, c.high_res_out), debug_info, debugger

  return agent_generator
