"""FEN codec (host mirror + oracle) against the reference's own outputs and documented strings."""
import numpy as np
import pytest
import torch

from chess_vision_b200 import dataset
from oracle import square_oracle as oracle


def test_class_tables_match_reference(golden):
    _, meta = golden
    assert dataset.CLASS_TO_TYPE == meta["class_tables"]["type"] == oracle.CLASS_TO_TYPE
    assert dataset.CLASS_TO_COLOR == meta["class_tables"]["color"] == oracle.CLASS_TO_COLOR
    assert "".join(dataset.INDEX_TO_PIECE[i] for i in range(13)) == meta["class_tables"]["pieces"] == oracle.PIECES
    assert dataset.NUM_CLASSES == 13 and dataset.NUM_SQUARES == 64


def test_known_answers_roundtrip(golden):
    _, meta = golden
    for fen, back in meta["kat_roundtrip"].items():
        assert back == fen                                        # the reference itself round-trips
        labels = dataset.fen_to_labels(fen)
        assert labels.tolist() == meta["kat_labels"][fen]
        assert dataset.labels_to_fen(labels) == fen
        assert oracle.placement_from_classes(labels.tolist()) == fen


def test_readme_example():
    # README.md:116 of the reference
    full = "rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq"
    p = dataset.parse_full_fen(full)
    assert dataset.labels_to_fen(p["squares"]) == full.split()[0]
    assert p["turn"].tolist() == [1.0] and p["castling"].tolist() == [1.0, 1.0, 1.0, 1.0]
    assert dataset.assemble_fen(full.split()[0], 0.3, [1, 1, 1, 1]) == full


def test_parse_full_fen_matches_reference(golden):
    _, meta = golden
    p = dataset.parse_full_fen("rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq -")
    assert p["turn"].tolist() == meta["parse_full_kat"]["turn"]
    assert p["castling"].tolist() == meta["parse_full_kat"]["castling"]
    assert p["squares"].tolist() == meta["parse_full_kat"]["squares"]
    q = dataset.parse_full_fen("8/8/8/8/8/8/8/8")
    assert q["turn"].tolist() == [0.0] and q["castling"].tolist() == [0.0] * 4


def test_filename_to_fen(golden):
    _, meta = golden
    assert dataset.filename_to_fen("1B1B1K2-3p1N2-8-8-8-8-8-1B6.jpeg") == meta["filename_kat"]


def test_random_labels_against_reference(golden):
    arrays, meta = golden
    for labels, want in zip(arrays["rand_labels"], meta["rand_labels_fen"]):
        assert dataset.labels_to_fen(torch.from_numpy(labels.astype(np.int64))) == want
        assert oracle.placement_from_classes(labels) == want


def test_errors_like_reference():
    with pytest.raises(AssertionError):
        dataset.fen_to_labels("8/8/8/8/8/8/8/7")          # 63 squares (dataset.py:48)
    with pytest.raises(KeyError):
        dataset.fen_to_labels("8/8/8/8/8/8/8/7x")


def test_oracle_fen_strings_edge_cases():
    sq = np.zeros((3, 64, 13), np.float32)                 # all-zero logits -> argmax 0 everywhere (H7)
    sq[1, :, 12] = 1.0                                     # all black kings
    sq[2, 0, 6] = 2.0; sq[2, 63, 1] = 0.5
    fens = oracle.fen_strings(sq, [0.0, 1.0, -1.0], [[0, 0, 0, 0], [1, 1, 1, 1], [-1, 2, -1, 3]])
    assert fens[0] == "8/8/8/8/8/8/8/8 w -"
    assert fens[1] == "/".join(["kkkkkkkk"] * 8) + " b KQkq"
    assert fens[2] == "K7/8/8/8/8/8/8/7P w Qq"
    flipped = oracle.fen_strings(sq[2:], [-1.0], [[-1, 2, -1, 3]], flipped=[1])
    assert flipped[0] == "P7/8/8/8/8/8/8/7K w Qq"
