from .eng import num_vocab, sequence_to_text, symbols, text_to_sequence  # noqa: F401
