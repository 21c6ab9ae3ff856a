"""chess_vision_b200: B200-native (sm_100a) implementation of cloudui/chess-vision's ChessSquareCNN
inference hot path behind the reference's ``build_model`` / ``predict`` surface.

    from chess_vision_b200 import build_model, predict
    model = build_model(cfg).to("cuda").eval();  model.load_state_dict(ckpt["model"])
    out = model(images)                           # {"squares","turn","castling"}  (reference contract)
    fens = model.predict_fen(uint8_boards)        # fused fast path -> ["rnbqkbnr/... w KQkq", ...]
"""
from .dataset import (CLASS_TO_COLOR, CLASS_TO_TYPE, INDEX_TO_PIECE, NUM_CLASSES, NUM_SQUARES, PIECE_TO_INDEX,
                      fen_to_labels, labels_to_fen, parse_full_fen)
from .models import ChessSquareCNN, build_model, build_square
from .predict import fen_from_outputs, get_transform, predict
from .preprocess import predict_images, resize_boards

__all__ = ["build_model", "build_square", "ChessSquareCNN", "predict", "fen_from_outputs", "get_transform", "resize_boards", "predict_images",
           "labels_to_fen", "fen_to_labels", "parse_full_fen", "PIECE_TO_INDEX", "INDEX_TO_PIECE",
           "NUM_CLASSES", "NUM_SQUARES", "CLASS_TO_TYPE", "CLASS_TO_COLOR"]
