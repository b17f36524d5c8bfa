from .quants import (quantize_q40, dequantize_q40, quantize_q80, dequantize_q80, quantize, dequantize,
                     F_32, F_16, F_Q40, F_Q80, tensor_bytes, parse_float_type, float_type_name)
from .model_file import ModelFile, write_model_header, write_tensor, HEADER_KEYS
from .tokenizer_file import write_tokenizer, read_tokenizer
