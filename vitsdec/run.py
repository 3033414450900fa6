import importlib
import sys

main = importlib.import_module("personalized_text-to-speech_b200.run").main

if __name__ == "__main__":
    sys.exit(main())
