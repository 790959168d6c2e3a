from .model_inference import calc_text_img_similarity, evaluate_prototype_shap_imp  # noqa: F401
