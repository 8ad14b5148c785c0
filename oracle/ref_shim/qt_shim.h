// qt_shim.h -- the sliver of Qt 5 the reference's hot-path translation units name, so that they compile
// UNMODIFIED with g++.  Ours, not the reference's.  No event loop, no GUI: message boxes print to stderr.
#pragma once
#include <stdio.h>
#include <unistd.h>
#include <string>

#define Q_OBJECT
#define signals public
#define slots
#define emit

class QString {
public:
    QString() {}
    QString(const char *s) : s_(s ? s : "") {}
    QString(const std::string &s) : s_(s) {}
    static QString number(int v) { return QString(std::to_string(v)); }
    std::string toStdString() const { return s_; }
    QString &operator+=(const QString &o) { s_ += o.s_; return *this; }
    friend QString operator+(const QString &a, const QString &b) { return QString(a.s_ + b.s_); }
    friend QString operator+(const QString &a, const char *b) { return QString(a.s_ + b); }
    friend QString operator+(const char *a, const QString &b) { return QString(std::string(a) + b.s_); }
private:
    std::string s_;
};

class QObject {
public:
    explicit QObject(QObject * = 0) {}
    virtual ~QObject() {}
    static QString tr(const char *s) { return QString(s); }
};
class QWidget : public QObject { public: explicit QWidget(QWidget * = 0) {} };
class QMainWindow : public QWidget { public: explicit QMainWindow(QWidget * = 0) {} };
class QDialog : public QWidget { public: explicit QDialog(QWidget * = 0) {} };

class QMessageBox {
public:
    enum { Yes = 1, No = 2 };
    static int warning(void *, const QString &title, const QString &text, int = 0, int = 0)
    {
        fprintf(stderr, "[QMessageBox::warning] %s: %s\n", title.toStdString().c_str(), text.toStdString().c_str());
        return 0;
    }
};
class QFile {
public:
    static bool exists(const QString &p) { return access(p.toStdString().c_str(), F_OK) == 0; }
};
