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


def host_debug_info(c, net, states, filters, cfg, index=0):
  """debug_info of image `index` in the reference's structure (agent.py:131-137, filters.py:79-87) as host
  numpy values -- what sess.run hands to the debugger: every filter's regressed parameters (only the
  selected filter was APPLIED, but all heads were evaluated), its mask, the pdf and the selected id.
  c: PolicyNet.forward context.  A few small D2H copies; not on the training path."""
  O = c.O[index]                                                       # [n_filters, ostride] raw fc2 outputs
  masking = bool(getattr(cfg, "masking", False))
  fdi = []
  for j, f in enumerate(filters):
    n = f.get_num_filter_parameters()
    p = f.filter_param_regressor(O[j:j + 1, :n].contiguous())
    nm = f.get_num_mask_parameters()
    if masking or nm != 6:
      mask = f.get_mask(net[index:index + 1], O[j:j + 1, n:n + nm].contiguous())
    else:
      mask = torch.ones(1, 1, 1, 1)                                    # filters.py:113
    fdi.append({"filter_parameters": (p if f.debug_info_batched() else p[0]).detach().cpu().numpy(),
                "mask": mask[0].detach().cpu().numpy()})
  return {
      "state": states,                                                 # agent.py:133
      "selected_filter_id": int(c.ids[index]),                         # agent.py:134
      "filter_debug_info": fdi,                                        # agent.py:135
      "pdf": c.pdf[index].detach().cpu().numpy(),                      # agent.py:136
  }


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
    filters = [cls(net, cfg) for cls in cfg.filters]
    debug_info = host_debug_info(c, net, states, filters, cfg, 0)      # first image, like the reference
    # batch-wide extras (not in the reference's dict)
    debug_info.update(selected_filter_ids=c.ids, selected_filter_parameters=c.params, penalty=c.penalty)
    from .visualize import make_debugger
    debugger = make_debugger(filters, int(net.shape[1]))              # agent.py:141-204

    if high_res is None:
      return (c.out, c.new_states, c.surrogate.view(B, 1), c.penalty.view(B, 1)), debug_info, debugger
    return (c.out, c.new_states, c.high_res_out), debug_info, debugger

  return agent_generator
