// parser.cpp -- equation strings -> plan description (implicit monomials + grouped explicit terms).
//
// Grammar (kept byte-compatible with /root/reference/src/parser.cpp):
//   equation := lhs '=' rhs                       exactly one '='; spaces are irrelevant
//   side     := term (('+'|'-') term)*            split at parenthesis depth 0
//   term     := factor ('*' factor | '/' factor)* '/' means "* 1/"; "/(" is an error
//   factor   := number | parameter | 1/number | 1/parameter | q^n | iqx[^n] | iqy[^n] | iqz[^n] | 1/q[^n]
//             | field[^n] | '(' side ')'
// Parentheses are distributed until none are left; expanded terms are appended at the END of the term
// list (this fixes the order in which explicit terms are created).  RHS terms with the same sorted field
// multiset are merged into one term carrying several monomials.  LHS: "dt<field>" marks a dynamic field and
// flips the sign of the remaining (implicit) monomials; a lone "1*field" LHS is dropped.
#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "../../inc/cupss.h"

parser::parser(evolver *s) : system(s) {}

// ------------------------------------------------------------------ parameters
int parser::exists_parameter(const std::string &name) { return parameters.count(name) ? 1 : 0; }

int parser::insert_parameter(const std::string &name, float value) {
    if (exists_parameter(name)) {
        std::cout << "Duplicate parameter " << name << std::endl;
        return -1;
    }
    parameters[name] = value;
    return 0;
}

float parser::getParameter(const std::string &name) {
    if (!exists_parameter(name)) {
        std::cout << "ERROR: getParameter " << name << " not found" << std::endl;
        std::exit(1);
    }
    return parameters[name];
}

int parser::changeParameter(const std::string &name, float value) {
    if (!exists_parameter(name)) {
        std::cout << "ERROR: changeParameter " << name << " not found" << std::endl;
        std::exit(1);
    }
    parameters[name] = value;
    return 0;
}

void parser::writeParamsToFile(const std::string &path) {
    std::ofstream out(path);
    for (const auto &kv : parameters) out << kv.first << "\t" << kv.second << "\n";
}

// ------------------------------------------------------------------ string helpers
std::string parser::strip_spaces(const std::string &s) {
    std::string r;
    for (char c : s) if (c != ' ') r.push_back(c);
    return r;
}

void parser::split_sides(const std::string &eq, std::string &lhs, std::string &rhs) {
    if (std::count(eq.begin(), eq.end(), '=') != 1) {
        std::cout << "Error processing" << std::endl << eq << std::endl << "More than one equal sign" << std::endl;
        std::exit(1);
    }
    const size_t pos = eq.find('=');
    lhs = eq.substr(0, pos);
    rhs = eq.substr(pos + 1);
}

// "a-b*(c+d)+e" -> {"a", "-b*(c+d)", "+e"}; a sign at position 0 stays attached to the first term.
void parser::split_sum(const std::string &s, std::vector<std::string> &terms) {
    int depth = 0;
    size_t start = 0;
    for (size_t i = 0; i < s.size(); i++) {
        if (s[i] == '(') depth++;
        if (s[i] == ')') depth--;
        if (i > 0 && depth == 0 && (s[i] == '+' || s[i] == '-')) {
            terms.push_back(s.substr(start, i - start));
            start = i;
        }
    }
    terms.push_back(s.substr(start));
}

// Factors separated by '*' outside parentheses (one level of parentheses, like the reference).
void parser::split_product(const std::string &s, std::vector<std::string> &factors) {
    bool inside = false;
    size_t start = 0;
    for (size_t i = 0; i < s.size(); i++) {
        if (s[i] == '(') inside = true;
        if (s[i] == ')') inside = false;
        if (s[i] == '*' && !inside) {
            factors.push_back(s.substr(start, i - start));
            start = i + 1;
        }
        if (i + 1 == s.size()) factors.push_back(s.substr(start));
    }
}

// Distribute the first top-level parenthesis of `t`; returns false when there is none.
bool parser::distribute_once(const std::string &t, std::vector<std::string> &out) {
    int depth = 0;
    size_t open = 0;
    bool seen = false;
    for (size_t i = 0; i < t.size(); i++) {
        if (t[i] == '(') {
            if (i > 0 && t[i - 1] == '/') {
                std::cout << "ERROR: No dividing over parenthesis allowed!" << std::endl;
                std::exit(1);
            }
            seen = true;
            if (++depth == 1) open = i;
        }
        if (t[i] == ')') depth--;
        if (seen && depth == 0) {
            std::vector<std::string> inner;
            split_sum(t.substr(open + 1, i - open - 1), inner);
            const std::string head = t.substr(0, open), tail = t.substr(i + 1);
            for (std::string piece : inner) {
                std::string h = head;
                if (!piece.empty() && piece[0] == '+') piece.erase(0, 1);
                if (!piece.empty() && piece[0] == '-') {   // a minus inside flips the sign carried by the head
                    if (!h.empty() && h[0] == '-') h[0] = '+';
                    else if (!h.empty() && h[0] == '+') h[0] = '-';
                    else h = "-" + h;
                    piece.erase(0, 1);
                }
                out.push_back(h + piece + tail);
            }
            return true;
        }
    }
    return false;
}

void parser::distribute_all(std::vector<std::string> &terms) {
    for (size_t k = 0; k < terms.size();) {
        std::vector<std::string> expanded;
        if (distribute_once(terms[k], expanded)) {
            terms.erase(terms.begin() + k);
            terms.insert(terms.end(), expanded.begin(), expanded.end());   // appended at the end; slot k is re-examined
        } else {
            k++;
        }
    }
}

// "a/b" -> "a*1/b"
std::string parser::normalise_division(const std::string &t) {
    std::string r;
    for (char c : t) {
        if (c == '/') r += "*1";
        r.push_back(c);
    }
    return r;
}

// ------------------------------------------------------------------ factor classification
bool parser::looks_numeric(const std::string &text) {
    std::string s = text;
    if (s.compare(0, 2, "1/") == 0) s.erase(0, 2);
    if (std::count(s.begin(), s.end(), '.') > 1) {
        std::cout << "ERROR: prefactor not a parameter and not a number: " << text << std::endl;
        std::exit(1);
    }
    for (char c : s)
        if (c != '.' && !isdigit((unsigned char)c)) return false;
    return true;
}

// "name" -> 1, "name^n" -> n (n != 1), anything else -> 0.  ("name^1" is not recognised, as in the reference.)
int parser::field_power(const std::string &factor) {
    int power = 1;
    size_t caret = 0;
    for (size_t i = 0; i < factor.size(); i++)
        if (factor[i] == '^') { power = atoi(factor.substr(i + 1).c_str()); caret = i; }
    const std::string candidate = power != 1 ? factor.substr(0, caret) : factor;
    for (field *f : system->fields)
        if (f->name == candidate) return power;
    return 0;
}

int parser::q_power(const std::string &factor) {
    if (factor.size() < 3 || factor[0] != 'q' || factor[1] != '^') return 0;
    const std::string n = factor.substr(2);
    if (!looks_numeric(n)) {
        std::cout << "ERROR in parser: " << n << " power of q not a number, missing *?" << std::endl;
        std::exit(1);
    }
    return atoi(n.c_str());
}

int parser::tagged_power(const std::string &factor, const std::string &tag) {
    if (factor.size() < 3 || factor.compare(0, 3, tag) != 0) return 0;
    if (factor.size() == 3) return 1;
    const std::string n = factor.substr(4);
    if (!looks_numeric(n)) {
        std::cout << "ERROR in parser: " << n << " power of " << tag << " not a number, missing *?" << std::endl;
        std::exit(1);
    }
    return atoi(n.c_str());
}

float parser::numeric_value(const std::string &factor) {
    if (q_power(factor) || tagged_power(factor, "iqx") || tagged_power(factor, "iqy") || tagged_power(factor, "iqz") ||
        tagged_power(factor, "1/q") || field_power(factor))
        return 1.0f;
    const bool reciprocal = factor.compare(0, 2, "1/") == 0;
    const std::string body = reciprocal ? factor.substr(2) : factor;
    float v;
    if (exists_parameter(body)) v = parameters[body];
    else if (looks_numeric(body)) v = std::stof(body);
    else {
        std::cout << "ERROR, parameter not found and not a number: " << factor << std::endl;
        std::exit(1);
    }
    return reciprocal ? 1.0f / v : v;
}

std::string parser::field_of_factor(const std::string &factor) {
    if (field_power(factor) == 1) return factor;
    return factor.substr(0, factor.rfind('^'));
}

// ------------------------------------------------------------------ term analysis
pres parser::prefactor_of(const std::string &term_in) {
    std::string t = term_in;
    pres p = {1.0f, 0, 0, 0, 0, 0};
    if (!t.empty() && t[0] == '+') t.erase(0, 1);
    if (!t.empty() && t[0] == '-') { p.preFactor = -1.0f; t.erase(0, 1); }
    std::vector<std::string> factors;
    split_product(normalise_division(t), factors);
    for (const std::string &f : factors) {
        if (field_power(f) > 0) continue;
        p.preFactor *= numeric_value(f);
        p.q2n += q_power(f) / 2;
        p.iqx += tagged_power(f, "iqx");
        p.iqy += tagged_power(f, "iqy");
        p.iqz += tagged_power(f, "iqz");
        p.invq += tagged_power(f, "1/q");
    }
    return p;
}

void parser::fields_of(const std::string &term_in, std::vector<std::string> &out) {
    std::string t = term_in;
    if (!t.empty() && t[0] == '+') t.erase(0, 1);
    if (!t.empty() && t[0] == '-') t.erase(0, 1);
    std::vector<std::string> factors;
    split_product(normalise_division(t), factors);
    for (const std::string &f : factors) {
        const int n = field_power(f);
        for (int i = 0; i < n; i++) out.push_back(field_of_factor(f));
    }
}

int parser::count_fields(const std::string &term_in) {
    std::string t = term_in;
    if (!t.empty() && (t[0] == '+' || t[0] == '-')) t.erase(0, 1);
    std::vector<std::string> factors;
    split_product(t, factors);
    int n = 0;
    for (const std::string &f : factors) n += field_power(f);
    return n;
}

std::string parser::lhs_field(const std::string &term_in) {
    std::string t = term_in;
    if (!t.empty() && (t[0] == '+' || t[0] == '-')) t.erase(0, 1);
    std::vector<std::string> factors;
    split_product(t, factors);
    std::string name;
    for (const std::string &f : factors)
        if (field_power(f) > 0) name = f;
    return name;
}

int parser::isParameterInString(const std::string &term_in, const std::string &parameter) {
    std::string t = term_in;
    if (!t.empty() && (t[0] == '+' || t[0] == '-')) t.erase(0, 1);
    std::replace(t.begin(), t.end(), '/', '*');
    std::vector<std::string> factors;
    split_product(t, factors);
    return (int)std::count(factors.begin(), factors.end(), parameter);
}

int parser::recalculateImplicits(const std::vector<std::string> &strings, std::vector<pres> &out, int dynamic) {
    for (size_t i = 0; i < strings.size(); i++) {
        out[i] = prefactor_of(strings[i]);
        if (dynamic) out[i].preFactor *= -1.0f;
    }
    return 0;
}

// ------------------------------------------------------------------ noise amplitude: numbers/parameters, q^n, 1/q^n
pres parser::add_noise(const std::string &expression) {
    std::vector<std::string> factors;
    split_product(strip_spaces(expression), factors);
    pres p = {1.0f, 0, 0, 0, 0, 0};
    for (const std::string &f : factors) {
        p.preFactor *= numeric_value(f);
        p.q2n += q_power(f) / 2;
        p.invq += tagged_power(f, "1/q");
    }
    return p;
}

// ------------------------------------------------------------------ equations
int parser::add_equation(const std::string &equation_in) {
    const std::string equation = strip_spaces(equation_in);
    if (verbose) std::cout << "Processing\n" << equation << "\n";

    std::string lhs, rhs;
    split_sides(equation, lhs, rhs);
    std::vector<std::string> lhs_terms, rhs_terms;
    split_sum(lhs, lhs_terms);
    split_sum(rhs, rhs_terms);
    distribute_all(lhs_terms);
    distribute_all(rhs_terms);

    bool dynamic = false;
    std::string target;
    if (lhs_terms[0].compare(0, 2, "dt") == 0) {
        dynamic = true;
        target = lhs_terms[0].substr(2);
        if (!field_power(target)) {
            std::cout << "Field not found: " << target << "\n";
            std::exit(1);
        }
        if (field_power(target) > 1) {
            std::cout << "Nonlinearity in lhs: " << target << "\n";
            std::exit(1);
        }
        lhs_terms.erase(lhs_terms.begin());
    } else {
        if (count_fields(lhs_terms[0]) != 1) {
            std::cout << "ERROR: Nonlinear term in left hand side: " << count_fields(lhs_terms[0]) << std::endl;
            std::cout << lhs_terms[0] << std::endl;
            std::exit(1);
        }
        target = lhs_field(lhs_terms[0]);
    }
    field *F = system->fieldsMap[target];

    // implicit (LHS) monomials, all linear in the target field
    std::vector<pres> implicits;
    for (const std::string &t : lhs_terms) {
        std::vector<std::string> fl;
        fields_of(t, fl);
        if (fl.size() != 1) {
            std::cout << "Nonlinear term in lhs\n";
            std::exit(1);
        }
        if (fl[0] != target) {
            std::cout << "ERROR: " << target << " and " << fl[0] << " incompatible in lhs\n";
            std::exit(1);
        }
        pres p = prefactor_of(t);
        if (dynamic) p.preFactor *= -1.0f;
        implicits.push_back(p);
    }

    // explicit (RHS) terms grouped by their sorted field multiset
    std::vector<std::vector<std::string>> group_fields;
    std::vector<std::vector<pres>> group_pres;
    std::vector<std::vector<std::string>> group_text;
    for (const std::string &t : rhs_terms) {
        std::vector<std::string> fl;
        fields_of(t, fl);
        std::sort(fl.begin(), fl.end());
        const pres p = prefactor_of(t);
        bool merged = false;
        for (size_t g = 0; g < group_fields.size(); g++)
            if (group_fields[g] == fl) {
                group_pres[g].push_back(p);
                group_text[g].push_back(t);
                merged = true;
            }
        if (!merged) {
            group_fields.push_back(fl);
            group_pres.push_back({p});
            group_text.push_back({t});
        }
    }

    for (const auto &kv : parameters) F->usedParameters[kv.first] = 0;
    const bool trivial_lhs = implicits.size() == 1 && implicits[0].preFactor == 1.0f && implicits[0].q2n == 0 &&
                             implicits[0].iqx == 0 && implicits[0].iqy == 0 && implicits[0].iqz == 0 && implicits[0].invq == 0;
    if (!trivial_lhs)
        for (size_t i = 0; i < implicits.size(); i++) {
            F->implicit.push_back(implicits[i]);
            F->addImplicitString(lhs_terms[i]);
            for (const auto &kv : parameters)
                if (isParameterInString(lhs_terms[i], kv.first)) F->usedParameters[kv.first] = 1;
        }

    for (size_t g = 0; g < group_fields.size(); g++) {
        system->createTerm(target, group_pres[g], group_fields[g]);
        term *T = F->terms.back();
        T->setPrefactorString(group_text[g]);
        for (const auto &kv : parameters) T->usedParameters[kv.first] = 0;
        for (const std::string &txt : group_text[g])
            for (const auto &kv : parameters)
                if (isParameterInString(txt, kv.first)) T->usedParameters[kv.first] = 1;
    }
    return 0;
}

// ------------------------------------------------------------------ text-file front end
// Sections "Fields" (name dynamic output), "Parameters" (name value), "Equations" (one per line); '#' comments.
int parser::createFromFile(const std::string &path) {
    std::cout << "Trying to read system from " << path << std::endl;
    std::ifstream in(path);
    std::string line;
    int section = -1;
    std::vector<std::string> equations;
    auto header = [](const std::string &l, bool first) -> int {
        if (l.compare(0, 6, "Fields") == 0 || l.compare(0, 6, "fields") == 0) return 0;
        if (l.compare(0, 10, "Parameters") == 0 || l.compare(0, 10, "parameters") == 0) return 1;
        if (l.compare(0, 9, "Equations") == 0 || (first && l.compare(0, 9, "equations") == 0)) return 2;   // lower case only as first header
        return -1;
    };
    while (std::getline(in, line)) {
        std::cout << line << std::endl;
        if (line.empty() || line[0] == '#') continue;
        const int h = header(line, section == -1);
        if (section == -1 && h == -1) {
            std::cout << "ERROR: reading file " << path << std::endl;
            std::cout << "First line must be either fields, parameters, or equations, it is:" << std::endl << line << std::endl;
            return -1;
        }
        if (h != -1) { section = h; continue; }
        std::istringstream iss(line);
        if (section == 0) {
            std::string name;
            int dyn, outp;
            if (!(iss >> name >> dyn >> outp)) {
                std::cout << "Error reading field line: " << line << std::endl << "Must be: field_name dynamic_value output" << std::endl;
                return -1;
            }
            std::cout << "Creating field: " << name << ", dynamic: " << dyn << std::endl;
            system->createField(name, dyn);
            system->setOutputField(name, outp);
        } else if (section == 1) {
            std::string name;
            float value;
            if (!(iss >> name >> value)) {
                std::cout << "Error reading parameter line: " << line << std::endl << "Must be: param_name value" << std::endl;
                return -1;
            }
            std::cout << "Creating parameter: " << name << " = " << value << std::endl;
            insert_parameter(name, value);
        } else {
            equations.push_back(line);
        }
    }
    for (const std::string &e : equations) add_equation(e);
    return 0;
}
