"""Hyper-parameter surface of the reference (xmcgan/configs/coco_xmc.py:18-88). ml_collections is not installed in the
build image, so `ConfigDict` below is a minimal attribute-style stand-in; any object with the same attributes
(including a real ml_collections.ConfigDict) is accepted by the rest of the package."""


class ConfigDict(dict):
  """Attribute-style dict (subset of ml_collections.ConfigDict used by the hot path)."""

  def __getattr__(self, k):
    try:
      return self[k]
    except KeyError as e:
      raise AttributeError(k) from e

  def __setattr__(self, k, v):
    self[k] = v

  def copy_and_resolve_references(self):
    return ConfigDict(self)


_DEFAULTS = dict(
    # run control / data (unused by the hot path, kept so reference configs round-trip)
    seed=42, eval_num=30000, eval_avg_num=3, num_train_steps=-1, log_loss_every_steps=1000, eval_every_steps=1000,
    checkpoint_every_steps=5000, dataset="mscoco", coco_version="2014", data_dir="data/", return_text=False,
    return_filename=False, trial=0, show_num=64, shuffle_buffer_size=1000, train_shuffle=True, num_epochs=500,
    eval_batch_size=7,
    # optimiser (train_utils.py:181-186)
    beta1=0.5, beta2=0.999, d_lr=0.0004, g_lr=0.0001, polyak_decay=0.999, d_step_per_g_step=2,
    # model
    batch_norm_group_size=-1, dtype="bfloat16", image_size=128, batch_size=56, df_dim=96, gf_dim=96, z_dim=128,
    model_name="xmc", g_spectral_norm=False, d_spectral_norm=True, architecture="xmc_net", gamma_for_g=15,
    word_contrastive=True, sentence_contrastive=True, image_contrastive=True, pretrained_image_contrastive=True,
    cond_size=16,
)

_TEST_OVERRIDES = dict(batch_size=2, eval_batch_size=2, eval_num=2, eval_avg_num=1, num_train_steps=2,
                       log_loss_every_steps=1, eval_every_steps=1, checkpoint_every_steps=1, df_dim=16, gf_dim=16,
                       z_dim=8, show_num=4, num_epochs=1, shuffle_buffer_size=10)


def get_config():
  """coco_xmc.get_config (coco_xmc.py:18-68)."""
  return ConfigDict(_DEFAULTS)


def get_test_config():
  """coco_xmc.get_test_config (coco_xmc.py:71-88)."""
  c = get_config()
  c.update(_TEST_OVERRIDES)
  return c
