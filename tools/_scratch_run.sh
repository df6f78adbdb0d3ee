timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -k "odd_image" 2>&1 | tail -8
