"""Host-side helpers shared by the model modules (mirror of big_vision/models/common.py:24-92)."""
import logging

from big_vision_b200 import utils as u


def merge_params(loaded, inited, dont_load=(), match_dtype=False):
  """Makes `loaded` match the structure of `inited`; names matching a `dont_load` regex keep their
  init value (or may be missing on either side); any other mismatch raises with both key lists."""
  if inited is None:
    return loaded
  dont_load = u.check_and_compile_patterns(dont_load)

  def should_merge(name):
    return not any(pattern.fullmatch(name) for pattern in dont_load)

  loaded_flat = dict(u.tree_flatten_with_names(loaded)[0])
  inited_flat = dict(u.tree_flatten_with_names(inited)[0])
  merged = {}
  for name, init_val in inited_flat.items():
    if name in loaded_flat and should_merge(name):
      merged[name] = loaded_flat[name]
      if match_dtype:
        merged[name] = loaded_flat[name].astype(init_val.dtype)
    else:
      logging.info("Ignoring checkpoint and using init value for %s", name)
      merged[name] = init_val

  def pp(title, names, indent="  "):
    return f"{title}:\n" + "\n".join(f"{indent}{k}" for k in sorted(names)) if names else ""

  not_in_loaded = {k for k in inited_flat.keys() - loaded_flat.keys() if should_merge(k)}
  not_in_inited = {k for k in loaded_flat.keys() - inited_flat.keys() if should_merge(k)}
  if not_in_loaded or not_in_inited:
    raise ValueError(
        pp("Params in checkpoint", loaded_flat.keys()) + "\n" +
        pp("Params in model (code)", inited_flat.keys()) + "\n" +
        pp("Params in model (code) but not in checkpoint and not `dont_load`ed", not_in_loaded, indent=" - ") +
        "\n" +
        pp("Params in checkpoint but not in model (code) and not `dont_load`ed", not_in_inited, indent=" + "))
  return u.recover_tree(merged.keys(), merged.values())
