from vla_touch_b200.lstm_step_controller import TactileLSTMController, load_lstm_controller  # noqa: F401
