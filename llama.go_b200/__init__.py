"""llama.go_b200 — B200-native LLaMA forward-pass engine behind pkg/ml + pkg/llama.Eval's API."""
