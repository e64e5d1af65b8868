"""Model hyper-parameters of the configs named in BASELINE.json, as plain dicts (the ``model`` section of the
reference YAMLs; the drop-in also loads the YAML files themselves through ``config.load_yaml``).

Sources (reference repo): cfg/avmnist/avmnist_m2-mixer_{S,M,B}.yml:23-56, cfg/mimic/mimic_m2-mixer_H.yml:20-52,
cfg/mmimdb/mmimdb_3loss.yml:39-44 (pos_weight).  C4 (MM-IMDB synthetic) and C5 (Scaled) are the synthetic
definitions of SURVEY 8(d) - they do not exist in the reference.
"""
from __future__ import annotations

import copy


def _avmnist(hidden, token, channel, fusion_channel, n_enc, n_fus, dropout):
    return {
        "type": "AVMnistMixerMultiLoss", "dropout": dropout,
        "modalities": {
            "classification": {"num_classes": 10, "classifier": "StandardClassifier", "input_shape": [16, 49, hidden],
                               "hidden_dims": [1024, 512, 256, 32]},
            "image": {"block_type": "MLPMixer", "in_channels": 1, "hidden_dim": hidden, "patch_size": 14,
                      "image_size": [28, 28], "token_dim": token, "channel_dim": channel, "num_mixers": n_enc},
            "audio": {"block_type": "MLPMixer", "in_channels": 1, "hidden_dim": hidden, "patch_size": 56,
                      "image_size": [112, 112], "token_dim": token, "channel_dim": channel, "num_mixers": n_enc},
            "multimodal": {"block_type": "FusionMixer", "fusion_function": "ConcatFusion", "hidden_dim": hidden,
                           "token_dim": token, "channel_dim": fusion_channel, "num_mixers": n_fus},
        },
    }


AVMNIST_S = _avmnist(32, 16, 256, 256, 2, 1, 0.1)
AVMNIST_M = _avmnist(64, 16, 1024, 1024, 2, 1, 0.1)
AVMNIST_B = _avmnist(128, 32, 3072, 3078, 4, 2, 0.5)     # fusion channel_dim 3078 is what upstream ships
AVMNIST_OPTIM = {"lr": 1e-2, "betas": [0.9, 0.999], "eps": 1e-8, "weight_decay": 0.0, "scheduler_patience": 2}

MIMIC_H = {
    "type": "MimicMixerMultiLoss", "dropout": 0.3, "gradblend": False,
    "modalities": {
        "classification": {"num_classes": 6, "classifier": "StandardClassifier", "input_shape": [16, 1024, 64]},
        "time": {"block_type": "MLPMixerNoPatching", "in_channels": 1, "embedding_dim": 12, "proj_dim": 64,
                 "hidden_dim": 64, "num_patch": 24, "token_dim": 16, "channel_dim": 64, "num_mixers": 1},
        "static": {"block_type": "MLP", "in_channels": 1, "input_dim": 5, "hidden_dim": 64, "num_blocks": 2,
                   "output_dim": 64},
        "multimodal": {"block_type": "FusionMixer", "fusion_function": "ConcatFusion", "hidden_dim": 64, "token_dim": 8,
                       "channel_dim": 64, "num_mixers": 1},
    },
}
MIMIC_OPTIM = {"lr": 1e-3, "weight_decay": 0.0}

MMIMDB_POS_WEIGHT = [4.57642832, 7.38544978, 10.79846869, 13.23391421, 15.59020924, 18.62735849, 22.48861048,
                     25.21711367, 74.50943396, 31.31641554, 31.79549114, 32.90833333, 39.64859438, 56.90201729,
                     40.46106557, 58.24483776, 67.3890785, 84.92473118, 58.33087149, 62.68253968, 114.13294798,
                     141.54121864, 116.83431953]


def mmimdb(image, text, multimodal, dropout=0.0):
    return {"type": "MMIMDBMixerMultiLoss", "dropout": dropout, "pos_weight": MMIMDB_POS_WEIGHT,
            "modalities": {"classification": {"num_classes": 23, "classifier": "StandardClassifier",
                                              "input_shape": [1, 1, multimodal["hidden_dim"]]},
                           "image": image, "text": text, "multimodal": multimodal}}


# C4: SURVEY 8(d) synthetic MM-IMDB (224x224/p16 image, 512-token PNLP text, 23 labels)
MMIMDB_C4 = mmimdb(
    {"block_type": "MLPMixer", "in_channels": 3, "hidden_dim": 256, "patch_size": 16, "image_size": [224, 224],
     "token_dim": 16, "channel_dim": 512, "num_mixers": 2},
    {"block_type": "PNLPMixer", "max_seq_len": 512, "hidden_dim": 256, "num_mixers": 2, "mlp_hidden_dim": 512,
     "bottleneck_window_size": 2, "bottleneck_features_size": 256},
    {"block_type": "FusionMixer", "fusion_function": "ConcatFusion", "hidden_dim": 256, "token_dim": 16,
     "channel_dim": 512, "num_mixers": 2})

# the reduced-size MM-IMDB-shaped model behind tests/golden/mmimdb_tiny_b6.npz
MMIMDB_TINY = mmimdb(
    {"block_type": "MLPMixer", "in_channels": 3, "hidden_dim": 64, "patch_size": 16, "image_size": [64, 48],
     "token_dim": 16, "channel_dim": 96, "num_mixers": 1},
    {"block_type": "PNLPMixer", "max_seq_len": 24, "hidden_dim": 64, "num_mixers": 1, "mlp_hidden_dim": 48,
     "bottleneck_window_size": 1, "bottleneck_features_size": 40},
    {"block_type": "FusionMixer", "fusion_function": "ConcatFusion", "hidden_dim": 64, "token_dim": 16, "channel_dim": 96,
     "num_mixers": 1})


# C5: SURVEY 8(d) "Scaled M2-Mixer" (Mixer-B/16-width blocks x 12 per modality, 178.6 M parameters): two 3 x 224 x 224
# encoders through the AV-MNIST task module (dict keys 'image' / 'audio'), CE multi-head loss, K = 10.
def _scaled_encoder():
    return {"block_type": "MLPMixer", "in_channels": 3, "hidden_dim": 768, "patch_size": 16, "image_size": [224, 224],
            "token_dim": 384, "channel_dim": 3072, "num_mixers": 12}


SCALED_C5 = {
    "type": "AVMnistMixerMultiLoss", "dropout": 0.0,
    "modalities": {
        "classification": {"num_classes": 10, "classifier": "StandardClassifier", "input_shape": [16, 392, 768]},
        "image": _scaled_encoder(), "audio": _scaled_encoder(),
        "multimodal": {"block_type": "FusionMixer", "fusion_function": "ConcatFusion", "hidden_dim": 768, "token_dim": 384,
                       "channel_dim": 3072, "num_mixers": 12},
    },
}


def get(name: str) -> dict:
    return copy.deepcopy({"avmnist_S": AVMNIST_S, "avmnist_M": AVMNIST_M, "avmnist_B": AVMNIST_B, "mimic_H": MIMIC_H, "mmimdb_C4": MMIMDB_C4,
                          "mmimdb_tiny": MMIMDB_TINY, "scaled_C5": SCALED_C5}[name])
