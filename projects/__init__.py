"""Drop-in replacements for the files of the reference's ``projects`` package that sit on the decode hot path
(SURVEY.md section 8b).  Copy (or overlay) ``projects/models/UMGen.py``, ``projects/tokenizer/vq_model.py`` and
``projects/tools/decode_map.py`` over the reference's files; everything else of the reference tree (plugin/,
configs/, tools/evaluate.py, tools/model_pl.py ...) is used unchanged."""
