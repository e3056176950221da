// oracle/model_patch.hpp — TEST INFRASTRUCTURE ONLY (force-included by `make -C oracle refmodel`).
//
// The reference GPEngine hard-wires JC69 (`JC69Model substitution_model_;`, /root/reference/src/gp_engine.hpp:366)
// although everything it does with the model goes through the generic SubstitutionModel interface
// (GetEigenvectors / GetInverseEigenvectors / GetEigenvalues / GetFrequencies, gp_engine.hpp:367-376). To get
// reference outputs under GTR and HKY WITHOUT editing any reference file, the TUs that see
// GPEngine's layout are recompiled with this header force-included: it pulls in the reference's
// substitution_model.hpp first (so JC69Model itself is defined as usual), then renames the token JC69Model to a
// wrapper whose constructor builds the reference's own GTRModel / HKYModel / JC69Model from the environment:
//     BITO_REF_MODEL="GTR r_AC r_AG r_AT r_CG r_CT r_GT  pi_A pi_C pi_G pi_T"   (substitution_model.cpp:103-186)
//     BITO_REF_MODEL="HKY kappa  pi_A pi_C pi_G pi_T"                            (substitution_model.cpp:79-101, 33-77)
//     unset or "JC69"                                                            (the stock engine)
// The eigendecomposition, its ordering and its rounding are therefore the reference's.
#pragma once
#include <cstdlib>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "substitution_model.hpp"

class EnvSelectedSubstitutionModel {
 public:
  EnvSelectedSubstitutionModel() {
    const char* env = std::getenv("BITO_REF_MODEL");
    std::istringstream in(env != nullptr ? env : "JC69");
    std::string name;
    in >> name;
    std::vector<double> v;
    for (double x; in >> x;) v.push_back(x);
    if (name == "GTR") {
      if (v.size() != 10) Failwith("BITO_REF_MODEL=GTR needs 6 rates and 4 frequencies");
      model_ = std::make_unique<GTRModel>();
    } else if (name == "HKY") {
      if (v.size() != 5) Failwith("BITO_REF_MODEL=HKY needs kappa and 4 frequencies");
      model_ = std::make_unique<HKYModel>();
    } else {
      model_ = std::make_unique<JC69Model>();
    }
    if (!v.empty()) {
      // the parameter vector is laid out by the model's own block specification: fill it through the
      // named segments, as the reference's doctest does (substitution_model.hpp:127-140)
      EigenVectorXd params(static_cast<Eigen::Index>(v.size()));
      params.setZero();
      auto segments = model_->GetBlockSpecification().ParameterSegmentMapOf(params);
      auto rates = segments.at(SubstitutionModel::rates_key_);
      auto frequencies = segments.at(SubstitutionModel::frequencies_key_);
      const size_t n_rates = v.size() - 4;
      for (size_t i = 0; i < n_rates; ++i) rates[static_cast<Eigen::Index>(i)] = v[i];
      for (size_t i = 0; i < 4; ++i) frequencies[static_cast<Eigen::Index>(i)] = v[n_rates + i];
      model_->SetParameters(params);
    }
  }
  const EigenMatrixXd& GetEigenvectors() const { return model_->GetEigenvectors(); }
  const EigenMatrixXd& GetInverseEigenvectors() const { return model_->GetInverseEigenvectors(); }
  const EigenVectorXd& GetEigenvalues() const { return model_->GetEigenvalues(); }
  const EigenVectorXd& GetFrequencies() const { return model_->GetFrequencies(); }
  const EigenMatrixXd& GetQMatrix() const { return model_->GetQMatrix(); }

 private:
  std::unique_ptr<SubstitutionModel> model_;
};

#define JC69Model EnvSelectedSubstitutionModel
