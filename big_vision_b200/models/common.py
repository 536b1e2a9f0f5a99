"""Host-side helpers shared by the model modules.

`merge_params` has the contract of big_vision/models/common.py:24-92: the result has the structure
of the freshly initialised tree and the checkpoint's values, except for names matched by a
`dont_load` regex (those keep their init value and may be absent on either side); any other
structural difference is an error that lists both sides.
"""
from big_vision_b200 import utils as u


def _report(ckpt_names, model_names, only_model, only_ckpt):
  def block(title, names, bullet="  "):
    return [f"{title}:"] + [f"{bullet}{n}" for n in sorted(names)] if names else []
  lines = (block("Params in checkpoint", ckpt_names) + block("Params in model (code)", model_names) +
           block("Params in model (code) but not in checkpoint and not `dont_load`ed", only_model, " - ") +
           block("Params in checkpoint but not in model (code) and not `dont_load`ed", only_ckpt, " + "))
  return "\n".join(lines)


def merge_params(loaded, inited, dont_load=(), match_dtype=False):
  if inited is None:                 # nothing to match against (interactive use)
    return loaded
  patterns = u.check_and_compile_patterns(dont_load)
  exempt = lambda name: any(p.fullmatch(name) for p in patterns)
  ckpt = dict(u.tree_flatten_with_names(loaded)[0])
  model = dict(u.tree_flatten_with_names(inited)[0])

  out = {}
  for name, init_val in model.items():
    if name in ckpt and not exempt(name):
      out[name] = ckpt[name].astype(init_val.dtype) if match_dtype else ckpt[name]
    else:
      out[name] = init_val           # dont_load, or (checked below) missing from the checkpoint

  only_model = [n for n in model if n not in ckpt and not exempt(n)]
  only_ckpt = [n for n in ckpt if n not in model and not exempt(n)]
  if only_model or only_ckpt:
    raise ValueError(_report(ckpt.keys(), model.keys(), only_model, only_ckpt))
  return u.recover_tree(list(out.keys()), list(out.values()))
