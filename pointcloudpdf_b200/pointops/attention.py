"""attention_relation_step / attention_fusion_step (libs/pointops/functions/attention.py) are
exported by the reference package but never called by PTv1 or the recognizers (SURVEY.md 2.5,
section 8 f-4); they are declared here so ``from pointops import *`` keeps working and raise if used."""


def attention_relation_step(*args, **kwargs):
    raise NotImplementedError("pointops.attention_relation_step: outside the PTv1 hot path (SURVEY.md 8f)")


def attention_fusion_step(*args, **kwargs):
    raise NotImplementedError("pointops.attention_fusion_step: outside the PTv1 hot path (SURVEY.md 8f)")
