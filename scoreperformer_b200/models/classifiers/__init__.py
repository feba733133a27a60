from .model import (EmbeddingClassifierOutput, LinearEmbeddingClassifier, SequentialEmbeddingClassifier,
                    MultiHeadEmbeddingClassifier, MultiHeadEmbeddingClassifierConfig, MultiHeadEmbeddingClassifierOutput,
                    EmbeddingClassifiersRegistry)
